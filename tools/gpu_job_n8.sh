# One N-GPU box: the north-star whole-process run (config 4 from the text files through the drop-in CLI, all GPUs, both shardings),
# the cell-sharding probe (locus ranges x cell groups, raw NCCL all-reduce rate), then the bench at N GPUs (weak line + strong / cells
# sub-records).  usage: bash tools/gpu_job_n8.sh N [skip_cli]
set -x
N=${1:-8}
nvidia-smi -L | wc -l
if [ -z "$2" ]; then
  timeout 900 python tools/run_config4_cli.py --gpus $N --modes restarts,cells --out gpurun_out/config4_cli > gpurun_out/config4_cli_n$N.log 2>&1; echo "config4 cli rc=$?"; tail -9 gpurun_out/config4_cli_n$N.log | cut -c1-900
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/prof_cells.py > gpurun_out/prof_cells_n$N.json 2> gpurun_out/prof_cells_n$N.err; echo "prof rc=$?"; grep "^#" gpurun_out/prof_cells_n$N.err | cut -c1-300; tail -1 gpurun_out/prof_cells_n$N.json | python -c "import json,sys; print(json.loads(sys.stdin.read())['nccl_allreduce'])"
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.4g ms %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for k in ("strong","cells"):
    print(k, "%.4g" % d[k]["value"], "ms", d[k]["ms_per_step"], d[k].get("cell_groups"))
PY
