TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='Warning\|^\*\*\*\|OMP_NUM\|NCCL version'
echo "== dist_check p2p+graph"; timeout 300 $TR --master-port 29511 tools/dist_check.py 64 V 2>&1 | grep -v "$F" | tail -3 | cut -c1-400
echo "== dist_check F-cycle 128"; timeout 300 $TR --master-port 29515 tools/dist_check.py 128 F 2>&1 | grep -v "$F" | tail -3 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print({k:d[k] for k in ('shape','n_dist','efield_rel_diff')}, d['dist']['it_mg'], d['single']['it_mg'], d['dist']['wall_s'], d['single']['wall_s'])"
echo "== profile p2p"; timeout 300 $TR --master-port 29512 tools/dist_profile.py 256 3 2>&1 | grep -v "$F" | tail -3
echo "== profile p2p nograph"; EMG3D_B200_GRAPHS=0 timeout 300 $TR --master-port 29514 tools/dist_profile.py 256 3 2>&1 | grep -v "$F" | tail -2
