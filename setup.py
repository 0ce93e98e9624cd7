"""Build script of the drop-in `custma` package (replaces the reference's setup.py:28-57).

The reference builds one torch CUDAExtension with no arch flags.  Here `build_ext` compiles the hand-written
kernels with nvcc for sm_100a ONLY (-gencode arch=compute_100a,code=sm_100a) into the torch-free C-ABI library
custereomatching_b200/libcustma_b200.so, in-tree; the Python packages bind it with ctypes.

    python setup.py build_ext --inplace      # or: python -m custereomatching_b200.build
    pip install --no-build-isolation .
"""
import os
import sys

from setuptools import Command, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def get_version() -> str:
    scope = {}
    with open(os.path.join(ROOT, "custma", "version.py"), encoding="utf-8") as f:
        exec(compile(f.read(), "version.py", "exec"), scope)
    return scope["__version__"]


class build_ext(Command):
    description = "compile libcustma_b200.so with nvcc for sm_100a"
    user_options = [("inplace", "i", "accepted for compatibility; the library is always built in-tree"),
                    ("force", "f", "rebuild even if up to date")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = False
        self.force = False

    def finalize_options(self):
        pass

    def run(self):
        from custereomatching_b200.build import build, build_torch_module
        print(build(force=bool(self.force), verbose=True))
        print(build_torch_module(force=bool(self.force), verbose=True))   # custma/src.<abi>.so, the reference's module name


class build_py_with_lib(build_py):
    def run(self):
        self.run_command("build_ext")
        super().run()


setup(
    name="custma",
    version=get_version(),
    description="B200-native ZNCC stereo-matching cost volume (drop-in for lzhnb/CuStereoMatching's custma)",
    packages=["custma", "custereomatching_b200"],
    package_data={"custereomatching_b200": ["libcustma_b200.so", "csrc/*"], "custma": ["src.*.so"]},
    cmdclass={"build_ext": build_ext, "build_py": build_py_with_lib},
    zip_safe=False,
)
