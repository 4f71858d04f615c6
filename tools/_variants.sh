python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu --epsilon 0.08 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['recall_at_10'], d['config']['overflow_queries_per_step'])"
