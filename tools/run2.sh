TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== dist_check p2p"; timeout 300 $TR --master-port 29511 tools/dist_check.py 64 V 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM" | tail -4 | cut -c1-900
echo "== dist_check nccl"; EMG3D_B200_P2P=0 timeout 300 $TR --master-port 29513 tools/dist_check.py 64 V 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM" | tail -2 | cut -c1-900
echo "== profile p2p"; timeout 300 $TR --master-port 29512 tools/dist_profile.py 256 3 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM" | tail -3
echo "== profile nccl"; EMG3D_B200_P2P=0 timeout 300 $TR --master-port 29514 tools/dist_profile.py 256 3 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM" | tail -2
