O=gpurun_out/r2k; mkdir -p $O
for js in 0 1 2 3; do
  python tools/run_case.py 16384 32 1000 $js 0 0 fp32 3 >> $O/occ.txt 2>&1
  python tools/run_case.py 4096 64 1000 $js 0 0 fp32 3 >> $O/occ.txt 2>&1
done
grep -v "^.* [34][0-9]\.[0-9]* ms" $O/occ.txt
