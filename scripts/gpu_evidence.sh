#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, the bench line, the ncu launch list of the same
# bench command, and one `ncu --set full` capture of the blend / preprocess / sort / k-means kernels.
# Everything lands in gpurun_out/; scripts/summarize_profiles.py turns it into profiles/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-kmeans --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-kmeans --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# quick_bench launches 18 ogs:: kernels per frame; skip 4 frames, capture 2 whole frames
python scripts/quick_bench.py --iters 3 --prof 0 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"blend|preprocess|rs_|emit_kernel|scan_gather|ranges_kernel|set_scalar" -s 72 -c 36 \
    -o gpurun_out/${TAG}_raster python scripts/quick_bench.py --iters 3 --prof 0 > gpurun_out/${TAG}_ncu_raster.log 2>&1; echo "ncu raster rc=$?"
python scripts/kmeans_bench.py --iters 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kmeans_assign -s 3 -c 1 \
    -o gpurun_out/${TAG}_kmeans python scripts/kmeans_bench.py --iters 2 > gpurun_out/${TAG}_ncu_kmeans.log 2>&1; echo "ncu kmeans rc=$?"
# section-8f kernels (mask statistics, mask IoU, Adam, footprint votes): probes first, then one full capture of each
python scripts/stage3_bench.py > gpurun_out/${TAG}_stage3_probe.jsonl 2> gpurun_out/${TAG}_stage3_probe.err; echo "stage3 probe rc=$?"
python scripts/footprint_bench.py >> gpurun_out/${TAG}_stage3_probe.jsonl 2>> gpurun_out/${TAG}_stage3_probe.err; echo "footprint probe rc=$?"
python scripts/mask_bench.py >> gpurun_out/${TAG}_stage3_probe.jsonl 2>> gpurun_out/${TAG}_stage3_probe.err; echo "mask probe rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|mask_pack|mask_pair" -s 6 -c 3 \
    -o gpurun_out/${TAG}_stage3 python scripts/stage3_bench.py > gpurun_out/${TAG}_ncu_stage3.log 2>&1; echo "ncu stage3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"footprint_vote" -s 1 -c 1 \
    -o gpurun_out/${TAG}_footprint python scripts/footprint_bench.py > gpurun_out/${TAG}_ncu_footprint.log 2>&1; echo "ncu footprint rc=$?"
# GPU comparator (restatement of the upstream kernel structure): parity vs the product, then frames/s of both
python scripts/upstream_structure_bench.py > gpurun_out/${TAG}_upstream_structure.json 2> gpurun_out/${TAG}_upstream_structure.err; echo "comparator rc=$?"
# end-to-end Stage-1 loop (fused render + mask losses + backward + one-launch Adam): the loss must decrease
python scripts/stage1_train_demo.py > gpurun_out/${TAG}_stage1_demo.json 2> gpurun_out/${TAG}_stage1_demo.err; echo "stage1 demo rc=$?"
