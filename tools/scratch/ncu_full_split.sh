export SASSENA_SELF_PATH=split
ncu --set full --clock-control none --import-source on -k regex:self_split_fft -s 2 -c 1 -o gpurun_out/splitA_full -f python tools/scratch/p_split.py 10000 1024 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:self_split_combine -s 2 -c 1 -o gpurun_out/splitB_full -f python tools/scratch/p_split.py 10000 1024 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
