# GPU-box check of the conv kernels: parity tests, per-shape timings with/without the MMA tconv kernels, bench
python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo pytest rc=$?
for sh in "64 16 16 52 20 5" "64 32 32 26 20 5" "64 64 64 13 20 5" "1024 16 16 52 20 5" "1024 64 64 13 20 5"; do
  echo "== mma: $sh"; python scripts/profile_conv.py $sh
  echo "== old: $sh"; TAMGCN_DISABLE_TCONV_MMA=1 python scripts/profile_conv.py $sh
done > gpurun_out/s6_conv.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/s6_bench1.log 2>&1; echo bench rc=$?
TAMGCN_DISABLE_TCONV_MMA=1 python bench.py --no-cpu-baseline --no-roofline > gpurun_out/s6_bench1_old.log 2>&1; echo bench rc=$?
