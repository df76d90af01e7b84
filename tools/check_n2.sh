# 2-GPU sanity: distributed solve against the single-GPU solver (tile kernels on z-windows), bench line
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='Warning\|^\*\*\*\|OMP_NUM\|NCCL version'
echo "== dist_check V-cycle 128"; timeout 150 $TR --master-port 29517 tools/dist_check.py 128 V 2>&1 | grep -v "$F" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print({k:d[k] for k in ('shape','n_dist','lr','efield_rel_diff')}, 'it', d['dist']['it_mg'], d['single']['it_mg'], d['dist']['exit'])"
echo "== bench N=2"; timeout 150 $TR --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>&1 | grep -v "$F" | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 ms', d['ms_per_step'], 'value', d['value'], d['halo_exchange'])"
