#!/bin/bash
# Round-2 evidence on ONE GPU (under gpurun), in two parts so that each stays under gpurun's 64 MiB return limit:
#   part a: parity tests, the bench line, the ncu launch list of the same bench command, comparator, Stage-1 demo
#   part b: `ncu --set full` captures (one whole frame of the rasterizer; the two k-means Lloyd kernels)
#   part c: `ncu --set full` over one Stage-1 training step (BASELINE config 3)
#   part d: part a + the Stage-1 step in its three forms (eager / resident views / resident views + CUDA graph) with
#           the device timeline of the graphed step
#   part f: `ncu --set full` over one Stage-1 step on a RESIDENT view (12 kernels: feat_refresh, blend fwd/bwd, masks)
# Everything lands in gpurun_out/; scripts/summarize_profiles.py turns it into profiles/.
set -u
TAG=${1:-r2}
PART=${2:-a}
mkdir -p gpurun_out
if [ "$PART" = "a" ] || [ "$PART" = "d" ]; then
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --repeats 1 --no-kmeans --no-configs --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_bench_short.err; echo "short bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --repeats 1 --no-kmeans --no-configs --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/upstream_structure_bench.py > gpurun_out/${TAG}_upstream_structure.json 2> gpurun_out/${TAG}_upstream_structure.err; echo "comparator rc=$?"
python scripts/stage1_train_demo.py --lr 0.01 --iters 300 > gpurun_out/${TAG}_stage1_demo.json 2> gpurun_out/${TAG}_stage1_demo.err; echo "stage1 demo rc=$?"
python scripts/kmeans_seg_probe.py > gpurun_out/${TAG}_kmeans_probe.txt 2>&1; echo "kmeans probe rc=$?"
if [ "$PART" = "d" ]; then
rm -f gpurun_out/${TAG}_stage1_steps.txt
for m in "0 0" "1 0" "1 1"; do set -- $m; python scripts/stage1_steps.py --iters 60 --view-cache $1 --graph $2 >> gpurun_out/${TAG}_stage1_steps.txt 2>&1; done
cat gpurun_out/${TAG}_stage1_steps.txt
python scripts/stage1_graph_timeline.py > gpurun_out/${TAG}_stage1_graph_timeline.txt 2>&1; echo "timeline rc=$?"
fi
elif [ "$PART" = "f" ]; then
# 4 uncached steps of 27 library kernels, then 12 per step on the resident views: capture the second resident step
ncu --set full --clock-control none -k regex:"blend|preprocess|rs_|emit_kernel|scan_gather|ranges_kernel|set_scalar|mask_|cohesion|sam_ids|separation|feat_" -s 120 -c 12 \
    -o gpurun_out/${TAG}_stage1_cached python scripts/stage1_steps.py --iters 8 --view-cache 1 > gpurun_out/${TAG}_ncu_stage1_cached.log 2>&1; echo "ncu stage1 cached rc=$?"
ls -la gpurun_out/
elif [ "$PART" = "c" ]; then
python scripts/stage1_steps.py --iters 40 > gpurun_out/${TAG}_stage1_steps.txt; echo "stage1 steps rc=$?"; cat gpurun_out/${TAG}_stage1_steps.txt
# a step launches 27 kernels of the library: skip 3 steps, capture a window that holds one whole step
ncu --set full --clock-control none -k regex:"blend|preprocess|rs_|emit_kernel|scan_gather|ranges_kernel|set_scalar|mask_|cohesion|sam_ids|separation" -s 81 -c 30 \
    -o gpurun_out/${TAG}_stage1 python scripts/stage1_steps.py --iters 6 > gpurun_out/${TAG}_ncu_stage1.log 2>&1; echo "ncu stage1 rc=$?"
ls -la gpurun_out/
else
# quick_bench launches 18 ogs:: kernels per frame; skip 4 frames, capture ONE whole frame
python scripts/quick_bench.py --iters 3 --prof 0 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"blend|preprocess|rs_|emit_kernel|scan_gather|ranges_kernel|set_scalar" -s 72 -c 18 \
    -o gpurun_out/${TAG}_raster python scripts/quick_bench.py --iters 3 --prof 0 > gpurun_out/${TAG}_ncu_raster.log 2>&1; echo "ncu raster rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"blend_bwd" -s 4 -c 1 \
    -o gpurun_out/${TAG}_blend_bwd_src python scripts/quick_bench.py --iters 3 --prof 0 > gpurun_out/${TAG}_ncu_bwd.log 2>&1; echo "ncu blend_bwd rc=$?"
ncu --set full --clock-control none -k regex:"kmeans_assign" -s 54 -c 4 \
    -o gpurun_out/${TAG}_kmeans python scripts/kmeans_seg_probe.py > gpurun_out/${TAG}_ncu_kmeans.log 2>&1; echo "ncu kmeans rc=$?"
ls -la gpurun_out/
fi
