// Sliding-window forward: ZNCC cost volume + winner-take-all in O(1) work per cell (for the tiling see
// sliding_common.cuh).
//
// What it restates: forward_cost_volume_kernel (reference custma/src/stereo_matching_kernel.cu:17-72) and the WTA the
// examples run in torch (examples/verify.py:72-74).  The reference evaluates the k x k window of every cell from
// scratch (4k^2 loads per cell); here, for a fixed disparity s, the raw correlation sum(cam' * proj') over the window
// is a k x k box filter of the product image q_s[y,x] = cam'[y,x] * proj'[y,x-s]:
//   horizontal   k-term FMA chain for the first of a thread's 4 columns, then 2 FMAs per further column
//                (add the entering product, subtract the leaving one)
//   vertical     register ring of the last k-1 horizontal sums while the thread marches down the rows
//   epilogue     exy = box - A[h,w] * Sp[h,d]   (A = window mean of cam', Sp = window sum of proj', both pivoted)
//                cost = (exy + eps) * rsqrt(ex2[h,w] * ey2[h,d] + eps)                      (reference :71)
// The pivoted image rows and the per-pixel statistics of each row step arrive in an 8-slot shared-memory ring through
// cp.async + mbarrier (RowLoader, sliding_common.cuh), six steps ahead of their use; there is no __syncthreads in the
// row loop, and the only other global traffic is the volume itself (one 128-bit streaming store per 4 cells) and the
// WTA keys.
#include "sliding_common.cuh"

namespace custma {



template <int K, int NU, int WG, int MODE, int DIR, bool COST, bool WTA, bool HEAD>
__device__ __forceinline__ void forward_consumer(const Problem &p, float *smem, uint64_t *full_bar,
                                                 uint64_t *empty_bar, RowLoader<K, NU, WG> &loader, int b,
                                                 int h0, int rows, int w_base, int s_base, int steps,
                                                 float *__restrict__ cost, unsigned long long *__restrict__ wta_keys,
                                                 const HeadOut &head, int head_slot);


template <int K, int NU, int WG, bool COST, bool WTA, bool HEAD>
__global__ void __launch_bounds__(16 * NU * WG, 2)
    sliding_forward_kernel(const Problem p, const SlidingLayout L, const char *__restrict__ ws,
                           float *__restrict__ cost, unsigned long long *__restrict__ wta_keys, const uint32_t tc_threshold,
                           const HeadOut head) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    using G = SlideGeom<K, NU, WG>;
    constexpr int WTC = G::WTC, SC = G::SC, NS = G::NS, NCW = G::NCW;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];

    const int tid = threadIdx.x;
    const int wt = blockIdx.x / L.n_chunks, ch = blockIdx.x % L.n_chunks, nb = blockIdx.y, b = blockIdx.z;
    const int w_base = wt * WTC, s_base = chunk_s_base(L, p.W, w_base, ch), h0 = nb * L.RB;
    const int rows = min(L.RB, p.H - h0);
    // row steps, padded to whole periods of the pair-sum ring (the band copies and statistics rows cover the padding)
    const int steps = (rows + K - 1 + G::PERIOD - 1) / G::PERIOD * G::PERIOD;
    // so many flagged tiles that the tensor-core kernel (tc_forward.cu) computes the whole call
    if (*reinterpret_cast<const uint32_t *>(ws + L.off_fb_count) > tc_threshold) return;
    // ill-conditioned tiles belong to the direct two-pass kernels (sliding_fallback.cu)
    if (reinterpret_cast<const uint8_t *>(ws + L.off_flags)[tile_index(L, b, nb, wt, ch)]) return;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full_bar[i], G::NCONS);
            mbar_init(&empty_bar[i], NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();
    RowLoader<K, NU, WG> loader;
    loader.init(L, ws, b, nb, h0, w_base, s_base);
    for (int t = 0; t < kLookahead && t < steps; ++t) loader.issue(t, smem, full_bar, empty_bar);

    // Bodies: MODE 0 = every cell of the tile is valid (check-free, 128-bit stores); 1 = 128-bit stores with a per-thread
    // validity mask (tiles that touch w - s < 0 or the right image border); 2 = scalar, fully checked (reference-shaped
    // volume, D not a multiple of the chunk).  DIR = direction of the sliding horizontal sums (BoxRing::step): towards
    // the zero padding a tile can see, or no sliding when that could be on both sides.
    const bool vec = p.banded && (p.D & 3) == 0 && s_base + SC <= p.D;
    const bool clean_left = w_base > 0 && w_base - (s_base + SC - 1) - (K / 2 + 3) >= 0;   // no padding left of any chain
    const bool clean_right = w_base + WTC + K <= p.W;
#define CUSTMA_FWD_BODY(MODE, DIR)                                                                                        \
    forward_consumer<K, NU, WG, MODE, DIR, COST, WTA, HEAD>(p, smem, full_bar, empty_bar, loader, b, h0, rows, w_base,    \
                                                            s_base, steps, cost, wta_keys, head, ch * NU)
    if (vec && clean_left && w_base + WTC <= p.W) CUSTMA_FWD_BODY(0, 1);
    else if (vec && clean_right) CUSTMA_FWD_BODY(1, 2);
    else if (vec) CUSTMA_FWD_BODY(1, 0);
    else CUSTMA_FWD_BODY(2, 0);
#undef CUSTMA_FWD_BODY
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
    return __shfl_xor_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ unsigned long long max_u64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

template <int K, int NU, int WG, int MODE, int DIR, bool COST, bool WTA, bool HEAD>
__device__ __forceinline__ void forward_consumer(const Problem &p, float *smem, uint64_t *full_bar,
                                                 uint64_t *empty_bar, RowLoader<K, NU, WG> &loader, int b,
                                                 int h0, int rows, int w_base, int s_base, int steps,
                                                 float *__restrict__ cost, unsigned long long *__restrict__ wta_keys,
                                                 const HeadOut &head, int head_slot) {
    static_assert(!HEAD || WTA, "the head needs the unit maxima of the winner-take-all reduction");
    using G = SlideGeom<K, NU, WG>;
    constexpr int CL = G::CL, PL = G::PL, NS = G::NS, PERIOD = G::PERIOD;
    const int tid = threadIdx.x;
    const int l16 = tid & 15, u = tid >> 4, su = u % NU, wg = u / NU;
    const int w0 = w_base + 4 * wg, s0 = s_base + 64 * su + 4 * l16;
    const int pidx = 4 * (wg - 16 * su - l16 + 16 * NU - 1);
    const int C = p.C;
    // pixel (h0 - (K-1), w0): output row of step t is h0 + t - (K-1), so the running pointers start k-1 rows early
    const int64_t pix_start = ((int64_t)b * p.H + h0 - (K - 1)) * p.W + w0;
    float *out = COST ? cost + pix_start * C + (MODE == 2 ? 0 : s0) : nullptr;
    uint32_t cmask = 0;  // MODE 1: bit 4i+j = cell (w0+i, s0+j) exists and is valid; constant over the band's rows
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (w0 + i < p.W && w0 + i - (s0 + j) >= 0) cmask |= 1u << (4 * i + j);
    }
    unsigned long long *keys = WTA ? wta_keys + pix_start : nullptr;
    // head: this unit's slot of the partial (m, z, n) image; same running row pointer as the keys
    float4 *hpart = HEAD ? head.part + (int64_t)(head_slot + su) * p.pixels() + pix_start : nullptr;
    const int64_t out_row = (int64_t)p.W * C;
    const float seed = kEps / (float)K;

    BoxRing2<K> ring;   // two disparities per FFMA2 / FADD2 (sliding_common.cuh): 7 % fewer instructions, same bits
    ring.clear();

#pragma unroll 1
    for (int t0 = 0; t0 < steps; t0 += PERIOD) {
#pragma unroll
        for (int q = 0; q < PERIOD; ++q) {
            const int t = t0 + q;
            const int slot = t & (NS - 1);
            if (t + kLookahead < steps) loader.issue(t + kLookahead, smem, full_bar, empty_bar);
            mbar_wait(&full_bar[slot], (t / NS) & 1);
            const float *S = smem + slot * G::SLOT;
            float c[CL], pj[PL];
#pragma unroll
            for (int v = 0; v < CL / 4; ++v)
                *reinterpret_cast<float4 *>(&c[4 * v]) = *reinterpret_cast<const float4 *>(S + 4 * wg + 4 * v);
#pragma unroll
            for (int v = 0; v < PL / 4; ++v)
                *reinterpret_cast<float4 *>(&pj[4 * v]) = *reinterpret_cast<const float4 *>(S + G::OFF_PROJ + pidx + 4 * v);
            f32x2 bx2[4][2];
            ring.template step<DIR>(q, c, pj, seed, bx2);
            if (t >= K - 1) {  // the first k-1 steps only fill the ring
                float a4[4], e4[4], sp[8], ey[8];
                *reinterpret_cast<float4 *>(a4) = *reinterpret_cast<const float4 *>(S + G::OFF_A + 4 * wg);
                *reinterpret_cast<float4 *>(e4) = *reinterpret_cast<const float4 *>(S + G::OFF_EX2 + 4 * wg);
                *reinterpret_cast<float4 *>(&sp[0]) = *reinterpret_cast<const float4 *>(S + G::OFF_SP + pidx);
                *reinterpret_cast<float4 *>(&sp[4]) = *reinterpret_cast<const float4 *>(S + G::OFF_SP + pidx + 4);
                *reinterpret_cast<float4 *>(&ey[0]) = *reinterpret_cast<const float4 *>(S + G::OFF_EY2 + pidx);
                *reinterpret_cast<float4 *>(&ey[4]) = *reinterpret_cast<const float4 *>(S + G::OFF_EY2 + pidx + 4);
                const bool row_ok = t - (K - 1) < rows;   // false only in the padding steps of a short last band
                unsigned long long key[4];
                float hv[HEAD ? 4 : 1][4];   // head: the costs of the thread's cells, -inf where no cell exists
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v[4];
                    float bv = -INFINITY;
                    int bs = 0;
                    float vals[4];
#pragma unroll
                    for (int jp = 0; jp < 2; ++jp) {   // disparities (2 jp, 2 jp + 1) in one instruction
                        const int di = i - 2 * jp + 3;  // projector column w_i - s_j, relative to w0 - s0 - 3 (j + 1: di - 1)
                        const f32x2 exy = fma2(pk2(-a4[i], -a4[i]), pk2(sp[di], sp[di - 1]), bx2[i][jp]);   // box already holds + eps
                        const f32x2 d = fma2(pk2(e4[i], e4[i]), pk2(ey[di], ey[di - 1]), pk2(kEps, kEps));
                        const f32x2 val2 = mul2(exy, pk2(rsqrt_fast(lo2(d)), rsqrt_fast(hi2(d))));          // reference kernel.cu:71
                        vals[2 * jp] = lo2(val2);
                        vals[2 * jp + 1] = hi2(val2);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float val = vals[j];
                        if (MODE == 0) {
                            v[j] = val;
                            if (HEAD) hv[i][j] = val;
                            if (WTA && (j == 0 || val >= bv)) { bv = val; bs = s0 + j; }
                        } else {
                            bool valid;
                            if (MODE == 1) {
                                valid = (cmask >> (4 * i + j)) & 1u;
                            } else {
                                const int d = w0 + i - (s0 + j);
                                valid = d >= 0 && d < p.W && (!p.banded || s0 + j < p.D);
                            }
                            v[j] = valid ? val : kInvalid;
                            if (HEAD) hv[i][j] = valid ? val : -INFINITY;
                            if (WTA && valid && val >= bv) { bv = val; bs = s0 + j; }
                        }
                    }
                    if (COST && row_ok) {
                        if (MODE == 0) {
                            __stcs(reinterpret_cast<float4 *>(out + (int64_t)i * C), make_float4(v[0], v[1], v[2], v[3]));
                        } else if (MODE == 1) {
                            if (w0 + i < p.W)
                                __stcs(reinterpret_cast<float4 *>(out + (int64_t)i * C), make_float4(v[0], v[1], v[2], v[3]));
                        } else if (w0 + i < p.W) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int s = s0 + j, d = w0 + i - s;
                                if (p.banded) {
                                    if (s < p.D) __stcs(out + (int64_t)i * C + s, v[j]);
                                } else if (d >= 0 && d < p.W) {
                                    __stcs(out + (int64_t)i * C + d, v[j]);
                                }
                            }
                        }
                    }
                    if (WTA)
                        key[i] = (MODE == 0 || bv > -INFINITY)
                                     ? ((unsigned long long)float_to_ordered(bv) << 32) | (uint32_t)(bs + p.W)
                                     : 0ull;
                }
                if (WTA) {
                    // reduce the 4 keys over the 16 lanes of the unit (reduce-scatter: 5 exchanges instead of 16)
                    const bool hi8 = l16 & 8, hi4 = l16 & 4;
                    unsigned long long k0 = hi8 ? key[2] : key[0], k1 = hi8 ? key[3] : key[1];
                    k0 = max_u64(k0, shfl_xor_u64(hi8 ? key[0] : key[2], 8));
                    k1 = max_u64(k1, shfl_xor_u64(hi8 ? key[1] : key[3], 8));
                    unsigned long long kk = max_u64(hi4 ? k1 : k0, shfl_xor_u64(hi4 ? k0 : k1, 4));
                    kk = max_u64(kk, shfl_xor_u64(kk, 2));
                    kk = max_u64(kk, shfl_xor_u64(kk, 1));
                    const int i = (hi8 ? 2 : 0) + (hi4 ? 1 : 0);
                    if ((l16 & 3) == 0 && row_ok && (MODE == 0 || (kk != 0ull && w0 + i < p.W))) atomicMax(keys + i, kk);
                    if (HEAD) {
                        // softmax partials relative to the unit's own maximum of every column: lanes 4i .. 4i+3 of the
                        // unit hold column i's maximum after the reduce-scatter above
                        const float kmax = kk != 0ull ? ordered_to_float((uint32_t)(kk >> 32)) : -INFINITY;
                        float z[4], n[4];
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float m = __shfl_sync(0xffffffffu, kmax, (tid & 16) | (4 * c4));
                            const float off = -head.beta_log2e * m;
                            float zz = 0.f, nn = 0.f;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                // exp2(-inf) = 0 for cells that do not exist; a unit without any cell has m = -inf and
                                // only such cells (the select keeps inf - inf out of the sums)
                                const float ex = hv[c4][j] > -INFINITY ? exp2_fast(fmaf(hv[c4][j], head.beta_log2e, off)) : 0.f;
                                zz += ex;
                                nn = fmaf(ex, (float)(s0 + j), nn);
                            }
                            z[c4] = zz;
                            n[c4] = nn;
                        }
                        // sum over the 16 lanes, same reduce-scatter (fixed order: deterministic)
                        float a0 = hi8 ? z[2] : z[0], a1 = hi8 ? z[3] : z[1], b0 = hi8 ? n[2] : n[0], b1 = hi8 ? n[3] : n[1];
                        a0 += __shfl_xor_sync(0xffffffffu, hi8 ? z[0] : z[2], 8);
                        a1 += __shfl_xor_sync(0xffffffffu, hi8 ? z[1] : z[3], 8);
                        b0 += __shfl_xor_sync(0xffffffffu, hi8 ? n[0] : n[2], 8);
                        b1 += __shfl_xor_sync(0xffffffffu, hi8 ? n[1] : n[3], 8);
                        float zs = (hi4 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi4 ? a0 : a1, 4);
                        float ns = (hi4 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, hi4 ? b0 : b1, 4);
                        zs += __shfl_xor_sync(0xffffffffu, zs, 2);
                        ns += __shfl_xor_sync(0xffffffffu, ns, 2);
                        zs += __shfl_xor_sync(0xffffffffu, zs, 1);
                        ns += __shfl_xor_sync(0xffffffffu, ns, 1);
                        if ((l16 & 3) == 0 && row_ok && (MODE == 0 || w0 + i < p.W))
                            hpart[i] = make_float4(kmax, zs, ns, 0.f);
                    }
                }
            }
            if (COST) out += out_row;
            if (WTA) keys += p.W;
            if (HEAD) hpart += p.W;
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty_bar[slot]);
        }
    }
}

// best / index from the packed keys, and the example-level outputs that depend on nothing else: the confidence mask
// best > threshold (examples/verify.py:74) and the masked disparity (column - correspondence) * mask
// (examples/test.py:83-84) - s already is column - correspondence
__global__ void __launch_bounds__(256)
    wta_decode_kernel(Problem p, const unsigned long long *__restrict__ keys, float *__restrict__ best,
                      int32_t *__restrict__ index, const WtaExtras ex) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.pixels()) return;
    const unsigned long long key = keys[pix];
    const int s = (int)(uint32_t)(key & 0xffffffffu) - p.W;
    const float bv = ordered_to_float((uint32_t)(key >> 32));
    best[pix] = bv;
    index[pix] = p.banded ? s : (int)(pix % p.W) - s;
    const float m = bv > ex.threshold ? 1.f : 0.f;
    if (ex.mask) ex.mask[pix] = m;
    if (ex.masked_disparity) ex.masked_disparity[pix] = (float)s * m;
}

// the same two outputs for the paths that write best / index themselves (direct kernels)
__global__ void __launch_bounds__(256)
    wta_extras_kernel(Problem p, const float *__restrict__ best, const int32_t *__restrict__ index, const WtaExtras ex) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.pixels()) return;
    const float m = best[pix] > ex.threshold ? 1.f : 0.f;
    const int idx = index[pix];
    if (ex.mask) ex.mask[pix] = m;
    if (ex.masked_disparity) ex.masked_disparity[pix] = (float)(p.banded ? idx : (int)(pix % p.W) - idx) * m;
}

// Merges the head's per-slot partials of a pixel: M = max m, Z = sum z * exp(beta (m - M)), N likewise; soft disparity
// N / Z (examples/verify.py:31-39 on the disparity axis; for the reference-shaped volume s = column - correspondence runs
// over every projector column, so N / Z = column - soft correspondence, examples/test.py:85).  Writes the masked soft
// disparity (test.py:86), best / index / mask as wta_decode_kernel does, and the state the backward needs.
__global__ void __launch_bounds__(256)
    head_decode_kernel(Problem p, const unsigned long long *__restrict__ keys, const HeadOut head,
                       float *__restrict__ soft, float *__restrict__ best, int32_t *__restrict__ index,
                       float *__restrict__ mask, float4 *__restrict__ state, float threshold) {
    pdl_wait();      // chained launch (common.cuh): the preceding kernel is complete and visible from here on
    pdl_release();
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.pixels()) return;
    float M = -INFINITY;
    for (int sl = 0; sl < head.slots; ++sl) M = fmaxf(M, head.part[(int64_t)sl * p.pixels() + pix].x);
    float Z = 0.f, N = 0.f;
    for (int sl = 0; sl < head.slots; ++sl) {
        const float4 q = head.part[(int64_t)sl * p.pixels() + pix];
        if (q.x > -INFINITY) {
            const float w = exp2_fast(head.beta_log2e * (q.x - M));
            Z = fmaf(q.y, w, Z);
            N = fmaf(q.z, w, N);
        }
    }
    const float sd = N / Z;                       // Z >= 1: the winning cell contributes exp(0)
    const float m = M > threshold ? 1.f : 0.f;
    soft[pix] = sd * m;
    if (mask) mask[pix] = m;
    if (best) {
        const unsigned long long key = keys[pix];
        const int s = (int)(uint32_t)(key & 0xffffffffu) - p.W;
        best[pix] = ordered_to_float((uint32_t)(key >> 32));
        index[pix] = p.banded ? s : (int)(pix % p.W) - s;
    }
    if (state) state[pix] = make_float4(head.beta_log2e * M, 1.f / Z, sd, m);
}

int launch_wta_extras(const Problem &p, const float *best, const int32_t *index, const WtaExtras &ex, cudaStream_t stream) {
    if (!ex.mask && !ex.masked_disparity) return CUSTMA_OK;
    wta_extras_kernel<<<(unsigned)((p.pixels() + 255) / 256), 256, 0, stream>>>(p, best, index, ex);
    CUSTMA_LAUNCH_CHECK("wta_extras_kernel");
    return CUSTMA_OK;
}

template <int K, int NU, int WG, bool COST, bool WTA, bool HEAD>
static int launch_one(const Problem &p, const SlidingLayout &L, const char *ws, float *cost,
                      unsigned long long *keys, uint32_t tc_threshold, const HeadOut &head, cudaStream_t stream) {
    const size_t smem = (size_t)kSlidingStages * SlideGeom<K, NU, WG>::SLOT * sizeof(float);
    const int threads = 16 * NU * WG;
    auto kern = sliding_forward_kernel<K, NU, WG, COST, WTA, HEAD>;
    CUSTMA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(L.n_wtiles * L.n_chunks, L.NB, p.B);
    CUSTMA_CUDA_CHECK(launch_chained(kern, grid, dim3(threads), smem, stream, p, L, ws, cost, keys, tc_threshold, head));
    CUSTMA_LAUNCH_CHECK("sliding_forward_kernel");
    return CUSTMA_OK;
}

template <int K, int NU, int WG>
static int launch_cfg(const Problem &p, const SlidingLayout &L, const char *ws, float *cost,
                      unsigned long long *keys, uint32_t thr, const HeadOut &head, cudaStream_t stream) {
    if (head.part) return launch_one<K, NU, WG, false, true, true>(p, L, ws, nullptr, keys, thr, head, stream);
    if (cost && keys) return launch_one<K, NU, WG, true, true, false>(p, L, ws, cost, keys, thr, head, stream);
    if (cost) return launch_one<K, NU, WG, true, false, false>(p, L, ws, cost, keys, thr, head, stream);
    return launch_one<K, NU, WG, false, true, false>(p, L, ws, cost, keys, thr, head, stream);
}

template <int K>
static int launch_k(const SlidingConfig &cfg, const Problem &p, const SlidingLayout &L, const char *ws, float *cost,
                    unsigned long long *keys, uint32_t thr, const HeadOut &head, cudaStream_t stream) {
    switch (cfg.NU) {
        case 1: return launch_cfg<K, 1, 12>(p, L, ws, cost, keys, thr, head, stream);
        case 2: return launch_cfg<K, 2, 6>(p, L, ws, cost, keys, thr, head, stream);
        case 3: return launch_cfg<K, 3, 4>(p, L, ws, cost, keys, thr, head, stream);
        default: return launch_cfg<K, 4, 3>(p, L, ws, cost, keys, thr, head, stream);
    }
}

bool sliding_forward_supported(const Problem &p) {
    SlidingConfig cfg;
    return sliding_pick_config(p, false, &cfg);
}

size_t sliding_forward_workspace_bytes(const Problem &p) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, false, &cfg)) return 0;
    SlidingLayout L;
    make_sliding_layout(p, cfg, false, &L);
    return L.total;
}

// Share of the fallback work list (flagged tiles x row groups) above which the whole call goes to the tensor-core
// kernel: the per-cell fallback runs at ~40 Gcell/s, the tensor-core kernel at ~430, the sliding kernel at ~850, so
// the cross-over is at about 5 % flagged.
constexpr double kTensorShare = 0.04;

int launch_sliding_forward(const Problem &p, const float *cam, const float *proj, float *cost, float *best,
                           int32_t *index, const WtaExtras &extras, void *workspace, size_t workspace_bytes,
                           bool force_tensor, cudaStream_t stream, CallPhase phase) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, false, &cfg)) return set_error(CUSTMA_ERR_UNSUPPORTED, "no sliding-window kernel for k=%d", p.k);
    SlidingLayout L;
    make_sliding_layout(p, cfg, false, &L);
    if (workspace_bytes < L.total)
        return set_error(CUSTMA_ERR_WORKSPACE, "sliding forward needs %zu workspace bytes, %zu given", L.total, workspace_bytes);
    char *ws = (char *)workspace;
    unsigned long long *keys = best ? (unsigned long long *)(ws + L.off_wta) : nullptr;
    int rc;
    if (force_tensor) {
        if (!tc_forward_supported(p))
            return set_error(CUSTMA_ERR_UNSUPPORTED, "CUSTMA_FLAG_TENSOR needs a banded volume with D %% 4 == 0, D <= 572 and k = 3 or 5");
        if (phase == kCallPrepareOnly) return CUSTMA_OK;   // the forced tensor-core kernels prepare nothing
        if ((rc = launch_tc_forward(p, cam, proj, cost, keys, nullptr, 0, stream))) return rc;   // writes every key itself
    } else {
        if (phase != kCallPrepared && (rc = launch_sliding_prep(p, L, cam, proj, ws, stream))) return rc;
        if (phase == kCallPrepareOnly) return CUSTMA_OK;
        const double items = (double)p.B * L.NB * L.n_wtiles * L.fb_groups;
        const uint32_t thr = tc_forward_supported(p) ? (uint32_t)(kTensorShare * items) : 0xffffffffu;
        rc = p.k == 3 ? launch_k<3>(cfg, p, L, ws, cost, keys, thr, HeadOut(), stream)
           : p.k == 5 ? launch_k<5>(cfg, p, L, ws, cost, keys, thr, HeadOut(), stream)
                      : launch_k<7>(cfg, p, L, ws, cost, keys, thr, HeadOut(), stream);
        if (rc) return rc;
        if ((rc = launch_fallback_forward(p, L, cam, proj, ws, cost, keys, HeadOut(), thr, stream))) return rc;
        if (thr != 0xffffffffu &&
            (rc = launch_tc_forward(p, cam, proj, cost, keys, (const uint32_t *)(ws + L.off_fb_count), thr, stream)))
            return rc;
    }
    if (best) {
        CUSTMA_CUDA_CHECK(launch_chained(wta_decode_kernel, dim3((unsigned)((p.pixels() + 255) / 256)), dim3(256), 0, stream, p,
                                         (const unsigned long long *)keys, best, index, extras));
        CUSTMA_LAUNCH_CHECK("wta_decode_kernel");
    }
    return CUSTMA_OK;
}

// ---- fused head ------------------------------------------------------------------------------------------------
static size_t head_part_bytes(const Problem &p, const SlidingLayout &L) {
    return align256((size_t)L.n_chunks * L.NU * p.pixels() * sizeof(float4));
}

size_t sliding_head_forward_workspace_bytes(const Problem &p) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, false, &cfg)) return 0;
    SlidingLayout L;
    make_sliding_layout(p, cfg, false, &L);
    return align256(L.total) + head_part_bytes(p, L);
}

int launch_sliding_forward_head(const Problem &p, const float *cam, const float *proj, float *soft_disparity, float *best,
                                int32_t *index, float *mask, float4 *head_state, float beta, float threshold,
                                void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    SlidingConfig cfg;
    if (!sliding_pick_config(p, false, &cfg))
        return set_error(CUSTMA_ERR_UNSUPPORTED, "the fused head needs a sliding-window kernel: kernel_size 3, 5 or 7, got %d", p.k);
    SlidingLayout L;
    make_sliding_layout(p, cfg, false, &L);
    const size_t need = align256(L.total) + head_part_bytes(p, L);
    if (workspace_bytes < need)
        return set_error(CUSTMA_ERR_WORKSPACE, "fused head forward needs %zu workspace bytes, %zu given", need, workspace_bytes);
    char *ws = (char *)workspace;
    unsigned long long *keys = (unsigned long long *)(ws + L.off_wta);
    HeadOut head;
    head.part = (float4 *)(ws + align256(L.total));
    head.slots = L.n_chunks * L.NU;
    head.beta_log2e = beta * 1.4426950408889634f;
    int rc;
    if ((rc = launch_sliding_prep(p, L, cam, proj, ws, stream))) return rc;
    // no hand-over to the tensor-core kernel here (it has no head epilogue): flagged tiles go to the per-cell fallback,
    // which writes the same partials
    const uint32_t thr = 0xffffffffu;
    rc = p.k == 3 ? launch_k<3>(cfg, p, L, ws, nullptr, keys, thr, head, stream)
       : p.k == 5 ? launch_k<5>(cfg, p, L, ws, nullptr, keys, thr, head, stream)
                  : launch_k<7>(cfg, p, L, ws, nullptr, keys, thr, head, stream);
    if (rc) return rc;
    if ((rc = launch_fallback_forward(p, L, cam, proj, ws, nullptr, keys, head, thr, stream))) return rc;
    CUSTMA_CUDA_CHECK(launch_chained(head_decode_kernel, dim3((unsigned)((p.pixels() + 255) / 256)), dim3(256), 0, stream, p,
                                     (const unsigned long long *)keys, head, soft_disparity, best, index, mask, head_state,
                                     threshold));
    CUSTMA_LAUNCH_CHECK("head_decode_kernel");
    return CUSTMA_OK;
}

}  // namespace custma
