# usage: run_ncu_pair.sh TAG  -- ncu --set full of the step-only and the headline (fused drive) kernel, V60 512^3
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:phys_chord -s 3 -c 1 -o gpurun_out/$1_step -f python scripts/ncu_chord.py > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:phys_chord -s 3 -c 1 -o gpurun_out/$1_seq -f python scripts/ncu_chord.py --drive > gpurun_out/n2.log 2>&1
