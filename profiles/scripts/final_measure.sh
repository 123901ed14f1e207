# end-of-round evidence run on one B200: TAG=r2k bash profiles/scripts/final_measure.sh   (outputs under gpurun_out/${TAG}_*)
set -x
TAG=${TAG:-r2k}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_reference.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_b.log 2>&1
ncu -f --set full --clock-control none --import-source on -k regex:"flow_photo_kernel|flow_photo_norm|flow_stencil|flow_loss_finalize" -s 8 -c 4 -o gpurun_out/${TAG}_step python profiles/scripts/one_step.py fused_step > gpurun_out/${TAG}_ncu.log 2>&1
for w in geom depth depth-texture depth-live; do ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_$w.csv python profiles/scripts/mode_step.py $w > /dev/null 2>&1; done
for tool in memcheck racecheck initcheck; do timeout 900 compute-sanitizer --tool $tool python tests/sanitizer_workload.py 2>&1 | grep -E "SUMMARY|workload done" | tail -3; done > gpurun_out/${TAG}_sanitizer.txt 2>&1
cat gpurun_out/${TAG}_sanitizer.txt
