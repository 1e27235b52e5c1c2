set -x
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests/test_sensor_schedule.py -m gpu -x -q > $O/t_sched.log 2>&1; echo "rc=$?" >> $O/t_sched.log
for cfg in "4096 64" "512 64" "1024 32"; do set -- $cfg; python tools/run_case.py $1 $2 1000 0 0 0 fp32 3 >> $O/cases.txt 2>&1; done
python -m pytest tests -m gpu -x -q > $O/t_gpu.log 2>&1; echo "rc=$?" >> $O/t_gpu.log
tail -30 $O/t_sched.log; cat $O/cases.txt; tail -5 $O/t_gpu.log
