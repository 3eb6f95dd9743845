# BASELINE config 4: imagenet_vitvq_large shapes, batch 256 over 8 GPUs (32 per GPU), one gradient all-reduce per step
O=gpurun_out/n8
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --config large --batch 32 --steps 5 --warmup 3 --extras "" > $O/large_b256_8gpu.json 2> $O/large_b256_8gpu.err
cut -c1-600 $O/large_b256_8gpu.json; tail -n 5 $O/large_b256_8gpu.err
