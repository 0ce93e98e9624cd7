for P in 1 2 3 4 6; do for m in single slots; do echo "P=$P $m: $(CUSTMA_HOST_PIPELINE=$m python tools/e2e_probe2.py $P lag1 2>&1 | tail -1)"; done; done
