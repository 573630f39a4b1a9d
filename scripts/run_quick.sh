timeout 900 python -m pytest tests -m gpu -q -k "roi or cascade" 2>&1 | tail -3
timeout 300 python scripts/bench_bwd.py 2>&1 | tail -1
