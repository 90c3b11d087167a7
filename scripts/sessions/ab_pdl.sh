timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do for f in 1 0; do DCGC_PDL=$f timeout 200 python bench.py --no-e2e --no-cpu-baseline --sub dmpnn 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('pdl=$f', d['value'], d['ms_per_step'], 'dmpnn', d['dmpnn']['ms_per_step'])"; done; done
