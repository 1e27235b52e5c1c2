O=gpurun_out/r2i; mkdir -p $O
python -m pytest tests/test_gpu_orca.py tests/test_gpu_dropin.py -m gpu -x -q > $O/t.log 2>&1; echo "rc=$?" >> $O/t.log
python tools/bench_configs.py c5 > $O/c5.jsonl 2> $O/c5.err
tail -5 $O/t.log; cat $O/c5.jsonl | cut -c1-300
