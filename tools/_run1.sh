timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for a in "10" "12" "13" "14" "15" "16" "17" "18" "20" "14 curve=2" "16 curve=2"; do echo "== $a"; python tools/one_msm.py $a 2>&1 | tail -1 | cut -c1-330; done
