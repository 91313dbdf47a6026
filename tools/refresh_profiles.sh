set -x
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_c2_n1.json 2> gpurun_out/final_bench_c2_n1.log || exit 1
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.log
B200_BENCH_PROFILE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_quad -s 2 -c 1 -f -o gpurun_out/final_scan_quad python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_full.log 2>&1
tail -3 gpurun_out/final_pytest_gpu.log
python -c "
import json;d=json.load(open('gpurun_out/final_bench_c2_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline'],d['parity_vs_oracle'],d['latency_batch1_ms_p50'],d['clocks'])"
cat gpurun_out/final_bench_reference.json | head -c 600
ls -la gpurun_out/final_*
