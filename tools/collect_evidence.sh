#!/bin/bash
# Collects the measurements behind profiles/r2_*: one bench line per BASELINE configuration (+ wide inputs, a sustained
# line, the reference arm), ncu launch lists and `--set full` captures of the main kernels.  Run on a B200:
#   bash tools/collect_evidence.sh            (writes gpurun_out/r2ev_*)
# Numbers printed by runs under ncu are never bench values; the bench lines come from the plain runs.
set -u
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; python bench.py "$@" > $O/r2ev_line_$name.json 2> $O/r2ev_line_$name.err; tail -c 200 $O/r2ev_line_$name.json; echo; }
run default
run half --workload half_512x1024_b8_n50
run crowd --workload crowd_1024x2048_b4_n500
run crowd_kmeans --workload crowd_1024x2048_b4_n500_kmeans
run coco --workload coco_800x1333_c80_n1000
run wide --inputs wide --no-cpu
run sustained --min-seconds 2.5 --no-e2e --no-cpu
run ring1 --ring 1 --no-e2e --no-cpu
ISG_SPLIT_KEEP=1 run split --no-e2e --no-cpu
ISG_SPLIT_KEEP=1 run split_ring1 --ring 1 --no-e2e --no-cpu
run ref_default --impl reference --steps 3
run ref_half --impl reference --steps 3 --workload half_512x1024_b8_n50
run ref_crowd --impl reference --steps 2 --workload crowd_1024x2048_b4_n500
python tests/aux_timings.py > $O/r2ev_aux_timings.txt 2>&1
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 150 --csv --log-file $O/r2ev_launches_default.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1 > $O/r2ev_ncu1.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file $O/r2ev_launches_coco.csv python bench.py --workload coco_800x1333_c80_n1000 --steps 2 --warmup 3 --no-e2e --no-cpu > $O/r2ev_ncu2.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/r2ev_launches_kmeans.csv python bench.py --workload crowd_1024x2048_b4_n500_kmeans --steps 2 --warmup 3 --no-e2e --no-cpu --ring 1 > $O/r2ev_ncu3.log 2>&1
full() { name=$1; kern=$2; skip=$3; shift 3; $NCU --set full --import-source on -k regex:$kern -s $skip -c 1 -f -o $O/r2ev_$name "$@" > $O/r2ev_ncu_$name.log 2>&1; }
full dense dense_v4 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
full dense_wide dense_v4 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1 --inputs wide
full topk_filter topk_filter 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
full nms_small nms_small 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
full polygons instance_polygons 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
full decode_boxes decode_boxes_kernel 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
ISG_SPLIT_KEEP=1 full dense_split dense_v4 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
ISG_SPLIT_KEEP=1 full topk_select_split topk_select 4 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1
ISG_SPLIT_KEEP=1 $NCU --metrics gpu__time_duration.sum -c 150 --csv --log-file $O/r2ev_launches_split.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --ring 1 > $O/r2ev_ncu4.log 2>&1
full kmeans kmeans_loop 2 python bench.py --workload crowd_1024x2048_b4_n500_kmeans --steps 2 --warmup 3 --no-e2e --no-cpu --ring 1
if [ -z "${EV_QUICK:-}" ]; then
full mask_area mask_area 2 python bench.py --workload coco_800x1333_c80_n1000 --steps 2 --warmup 3 --no-e2e --no-cpu
full mask_pair mask_pair 2 python bench.py --workload coco_800x1333_c80_n1000 --steps 2 --warmup 3 --no-e2e --no-cpu
fi
full nms_scan nms_scan 2 python bench.py --workload coco_800x1333_c80_n1000 --steps 2 --warmup 3 --no-e2e --no-cpu
if [ -z "${EV_QUICK:-}" ]; then
full fill_polygons fill_polygons 1 python tools/bench_fill.py
python tools/overlap_experiment.py > $O/r2ev_overlap.txt 2>&1
fi
python tools/bench_fill.py > $O/r2ev_fill_polygons.txt 2>&1
ls -la $O/r2ev_*.ncu-rep
