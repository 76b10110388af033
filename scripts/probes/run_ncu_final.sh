# usage: run_ncu_final.sh TAG -- ncu --set full of the step-only and the headline kernel + the launch list of the bench command
mkdir -p gpurun_out
bash scripts/probes/run_ncu_pair.sh $1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$1_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/$1_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/$1_step.ncu-rep "ncu --set full --clock-control none | step kernel alone (force from a field, no rho/u write-out) | V60 512^3 | B200" > gpurun_out/$1_step_summary.txt
python scripts/ncu_summary.py gpurun_out/$1_seq.ncu-rep "ncu --set full --clock-control none | bench.py headline: fused drive, rho/u written | V60 512^3 configs[2] | B200" > gpurun_out/$1_seq_summary.txt
cat gpurun_out/$1_step_summary.txt gpurun_out/$1_seq_summary.txt
