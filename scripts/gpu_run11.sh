cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C[23]"
