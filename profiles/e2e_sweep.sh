for parts in 4 5 6 8; do
  echo "parts $parts"; MLB_HOST_PARTS=$parts python bench.py --no-ess --no-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'])"
done
for ch in 4096 6144 8192 10240 12288; do
  echo "chunk $ch"; MLB_HOST_CHUNK=$ch python bench.py --no-ess --no-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'])"
done
