set -x
O=gpurun_out/r2final; mkdir -p $O
# (captured in an earlier call of this script; the reports exceed the 64 MiB a call may bring back together) python -m pytest tests -m gpu -x -q > $O/t_gpu.log 2>&1; echo "pytest rc=$?" >> $O/t_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 10 --warmup 3 --extra > $O/bench_n1.json 2> $O/bench_n1.err
python tools/bench_configs.py c2 c3 c4 c5 sensors > $O/configs.jsonl 2> $O/configs.err
python tools/bench_dropin.py > $O/dropin.jsonl 2> $O/dropin.err
bash tools/bench_sensor_plan.sh > $O/sensor_plan.txt 2>&1
for cfg in "2048 64" "1024 64" "512 64" "1024 32"; do set -- $cfg; python tools/run_case.py $1 $2 1000 0 0 0 fp32 3 >> $O/strong_cases.txt 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
# (captured in an earlier call of this script; the reports exceed the 64 MiB a call may bring back together) ncu --set full --clock-control none --import-source on -k regex:hsfm_warp_kernel -s 1 -c 1 -f -o $O/warp_c4 python tools/run_case.py 4096 64 1000 0 0 0 fp32 2 > $O/ncu_warp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsfm_duo_kernel -s 1 -c 1 -f -o $O/duo_512x64 python tools/run_case.py 512 64 1000 0 0 0 fp32 2 > $O/ncu_duo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsfm_duo_kernel -s 1 -c 1 -f -o $O/duo_1024x32 python tools/run_case.py 1024 32 1000 0 0 0 fp32 2 > $O/ncu_duo2.log 2>&1
# (captured in an earlier call of this script; the reports exceed the 64 MiB a call may bring back together) ncu --set full --clock-control none --import-source on -k regex:orca_kernel -s 1 -c 1 -f -o $O/orca_c5 python tools/bench_configs.py c5 > $O/ncu_orca.log 2>&1
tail -3 $O/t_gpu.log; head -c 600 $O/bench_n1.json; cat $O/strong_cases.txt; cat $O/dropin.jsonl | cut -c1-200
