for n in 10 12 14 15 16; do for L in 4 8 12 16; do echo "== $n chunk=$L"; python tools/one_msm.py $n chunk=$L 2>&1 | tail -1 | cut -c1-250; done; done
