for i in 1 2 3; do for layout in caller binned; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 10 --no-render --no-extras --no-cpu-baseline --layout $layout > gpurun_out/n2_$layout.json 2> gpurun_out/n2_$layout.err; echo "try $i $layout rc=$? $(head -c 120 gpurun_out/n2_$layout.json | cut -c40-120)"
done; done
