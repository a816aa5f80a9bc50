set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/b_batch.json 2> gpurun_out/b_batch.err; tail -3 gpurun_out/b_batch.err
python bench.py --workload giga --steps 3 --warmup 3 > gpurun_out/b_giga1.json 2> gpurun_out/b_giga1.err; tail -3 gpurun_out/b_giga1.err
python bench.py --workload streams1080p --steps 2 --warmup 3 --no-cpu > gpurun_out/b_streams1.json 2> gpurun_out/b_streams1.err; tail -3 gpurun_out/b_streams1.err
python bench.py --workload frame4k --steps 200 --warmup 10 --no-cpu > gpurun_out/b_4k.json 2> gpurun_out/b_4k.err; tail -3 gpurun_out/b_4k.err
cat gpurun_out/b_*.json | cut -c1-600
