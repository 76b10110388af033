set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:phys_chord -s 3 -c 1 -o gpurun_out/r02b_base_step -f python scripts/ncu_chord.py > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:phys_chord -s 3 -c 1 -o gpurun_out/r02b_base_seq -f python scripts/ncu_chord.py --drive > gpurun_out/n2.log 2>&1
tail -3 gpurun_out/n1.log gpurun_out/n2.log
