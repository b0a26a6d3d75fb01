#!/bin/bash
# Round-2 evidence pass on ONE GPU (under gpurun): parity tests, the bench line, the ncu launch list of the same
# bench command, one `ncu --set full` capture of two whole frames + the k-means passes, and the comparator's
# launch list.  Everything lands in gpurun_out/; scripts/summarize_profiles.py turns it into profiles/.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --repeats 1 --no-kmeans --no-configs --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --repeats 1 --no-kmeans --no-configs --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# quick_bench launches 18 ogs:: kernels per frame; skip 4 frames, capture 2 whole frames
python scripts/quick_bench.py --iters 3 --prof 0 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"blend|preprocess|rs_|emit_kernel|scan_gather|ranges_kernel|set_scalar" -s 72 -c 36 \
    -o gpurun_out/${TAG}_raster python scripts/quick_bench.py --iters 3 --prof 0 > gpurun_out/${TAG}_ncu_raster.log 2>&1; echo "ncu raster rc=$?"
python scripts/kmeans_seg_probe.py > gpurun_out/${TAG}_kmeans_probe.txt 2>&1; echo "kmeans probe rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"kmeans_assign" -s 54 -c 4 \
    -o gpurun_out/${TAG}_kmeans python scripts/kmeans_seg_probe.py > gpurun_out/${TAG}_ncu_kmeans.log 2>&1; echo "ncu kmeans rc=$?"
python scripts/upstream_structure_bench.py > gpurun_out/${TAG}_upstream_structure.json 2> gpurun_out/${TAG}_upstream_structure.err; echo "comparator rc=$?"
python scripts/stage1_train_demo.py --lr 0.01 --iters 300 > gpurun_out/${TAG}_stage1_demo.json 2> gpurun_out/${TAG}_stage1_demo.err; echo "stage1 demo rc=$?"
