# One 8-GPU box: the north-star whole-process run (config 4 from the text files, all GPUs, both shardings, vs one GPU), then the
# bench at N = 8 (weak line + strong / cells sub-records) and the config-5 restart sweep.
set -x
N=${1:-8}
nvidia-smi -L | wc -l
timeout 1500 python tools/run_config4_cli.py --gpus $N --out gpurun_out/config4_cli > gpurun_out/config4_cli_n$N.log 2>&1; echo "config4 cli rc=$?"; tail -12 gpurun_out/config4_cli_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config 5 --sweep > gpurun_out/sweep5_n$N.json 2> gpurun_out/sweep5_n$N.err; echo "sweep rc=$?"
cat gpurun_out/bench_n$N.json | head -c 3000
