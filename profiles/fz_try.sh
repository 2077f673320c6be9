cd /root/repo; mkdir -p gpurun_out/fz
for cfg in "85 85 100" "75 105 105" "75 105 100" "70 110 105" "80 100 100" "75 120 105"; do set -- $cfg
  echo "== small $1 w0 $2 w5 $3"
  NERF_FZ_SMALL=$1 NERF_FZ_W0=$2 NERF_FZ_W5=$3 python tests/fused_check.py 1024 192 10 2>&1 | grep "fused:"
done | tee gpurun_out/fz/weights.txt
