set -x
python -m pytest tests -m gpu -q > gpurun_out/r1j_tests.log 2>&1; echo tests rc=$?
python bench.py > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; echo bench rc=$?
python tools/sweep_crt.py --iters 10 > gpurun_out/r1j_sweep.jsonl 2> gpurun_out/r1j_sweep.err; echo sweep rc=$?
python tools/bench_she.py > gpurun_out/r1j_she.jsonl 2> gpurun_out/r1j_she.err; echo she rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1j_smoke.log 2>&1; echo smoke rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-m2 > gpurun_out/r1j_ncu_bench.log 2>&1; echo ncu1 rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches_tunnel.csv python tools/bench_she.py --cases tunnel --ciphertexts 2960 --iters 1 > gpurun_out/r1j_ncu_tunnel.log 2>&1; echo ncu2 rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1j_launches_mul1.csv python tools/_mul1.py > gpurun_out/r1j_ncu_mul1.log 2>&1; echo ncu3 rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_crt_mixed_mac -c 1 -o gpurun_out/r1j_mac python tools/bench_she.py --cases tunnel --ciphertexts 2960 --iters 1 > gpurun_out/r1j_ncu_mac.log 2>&1; echo ncu4 rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_mulks_pow2 -c 1 -o gpurun_out/r1j_mulks python tools/bench_she.py --cases mul --iters 1 > gpurun_out/r1j_ncu_mulks.log 2>&1; echo ncu5 rc=$?
