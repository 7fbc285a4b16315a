timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "vote or face" 2>&1 | tail -5
python tools/bench_vote.py 2>&1 | tail -3
python tools/bench_vote.py 32 2000 2>&1 | tail -3
