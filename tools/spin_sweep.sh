for t in 14 12 10; do for sp in 0 40 300; do B2K_HOST_SPIN_US=$sp THREADS=$t ITERS=24 python tools/e2e_iters.py 2>&1 | tail -1 | sed "s/^/threads $t spin_us $sp: /"; done; done
THREADS=-1 ITERS=24 python tools/e2e_iters.py 2>&1 | tail -2
