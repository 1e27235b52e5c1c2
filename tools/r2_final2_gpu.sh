# Final round-2 capture on one B200 (outputs: gpurun_out/r2final2/, summaries copied to profiles/ by hand):
#   bash tools/r2_final2_gpu.sh
set -x
O=gpurun_out/r2final2; mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 10 --warmup 3 --extra > $O/bench_n1.json 2> $O/bench_n1.err
python tools/bench_configs.py c2 c3 c4 c5 sensors > $O/configs.jsonl 2> $O/configs.err
python tools/bench_dropin.py > $O/dropin.jsonl 2> $O/dropin.err
bash tools/bench_sensor_plan.sh > $O/sensor_plan.txt 2>&1
for cfg in "2048 64" "1024 64" "512 64" "1024 32"; do set -- $cfg; python tools/run_case.py $1 $2 1000 0 0 0 fp32 3 >> $O/strong_cases.txt 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsfm_duo_kernel -s 1 -c 1 -f -o $O/duo_1024x64 python tools/run_case.py 1024 64 1000 0 0 0 fp32 2 > $O/ncu_duo3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsfm_duo_kernel -s 1 -c 1 -f -o $O/duo_2048x64 python tools/run_case.py 2048 64 1000 0 0 0 fp32 2 > $O/ncu_duo4.log 2>&1
head -c 400 $O/bench_n1.json; echo; cat $O/strong_cases.txt; cat $O/dropin.jsonl | cut -c1-120; ls -la $O
