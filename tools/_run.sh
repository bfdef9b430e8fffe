python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in c5 c5v c3; do timeout 300 python tools/soak.py $w 20000 2>&1 | tail -2; done
