# small-launch form of the fused trajectory kernel (mlb_duo.cuh): default dispatch against off
for n in 2048 8192 16384 32768 65536; do
  for duo in 0 -1; do
    echo -n "chains $n MLB_DUO_MAX_CHAINS=$duo: "
    if [ $duo = -1 ]; then unset MLB_DUO_MAX_CHAINS; else export MLB_DUO_MAX_CHAINS=$duo; fi
    python bench.py --n-chain $n --no-ess --no-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms %.4f steps/s %.4g e2e %.4g iters %.3f failed %.4f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['measured']['mean_projection_iters_per_step'], d['measured']['failed_chain_fraction']))"
  done
done
unset MLB_DUO_MAX_CHAINS
for wl in toy; do for n in 8192 65536; do for duo in 0 -1; do
  echo -n "$wl chains $n MLB_DUO_MAX_CHAINS=$duo: "
  if [ $duo = -1 ]; then unset MLB_DUO_MAX_CHAINS; else export MLB_DUO_MAX_CHAINS=$duo; fi
  python bench.py --workload $wl --n-chain $n --no-ess --no-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms %.4f steps/s %.4g e2e %.4g' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
done; done; done
