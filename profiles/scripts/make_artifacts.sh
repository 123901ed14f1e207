# after final_measure.sh came back: TAG=r2n bash profiles/scripts/make_artifacts.sh  -> profiles/${TAG}_* summaries + traffic.json
set -e
T=${TAG:-r2n}
ncu -i gpurun_out/${T}_step.ncu-rep --page raw --csv > /tmp/${T}.csv 2>/dev/null
python profiles/make_traffic.py "profiles/${T}_step_ncu_keys.txt (ncu --set full of one fused step: profiles/scripts/final_measure.sh)" < /tmp/${T}.csv > profiles/traffic.json
python profiles/ncu_keys.py < /tmp/${T}.csv > profiles/${T}_step_ncu_keys.txt
ncu -i gpurun_out/${T}_step.ncu-rep --page source --csv --kernel-name regex:flow_stencil 2>/dev/null | python profiles/phase_counts.py > profiles/${T}_stencil_phase_instruction_counts.txt
ncu -i gpurun_out/${T}_step.ncu-rep --page source --csv --kernel-name regex:flow_photo_kernel 2>/dev/null | python profiles/hot_lines.py 30 | awk '/^kernel 1/{exit} {print}' > profiles/${T}_photo_hot_lines.txt
for f in bench.json reference.json launches_ncu.csv sanitizer.txt; do cp gpurun_out/${T}_$f profiles/${T}_$f; done
python profiles/summarise_launches.py gpurun_out/${T}_launches_ncu.csv flow_photo_kernel > profiles/${T}_launch_summary.txt
for w in geom depth depth-texture depth-live; do n=$(echo $w | tr - _); python profiles/summarise_launches.py gpurun_out/${T}_launches_$w.csv pose_setup_fwd "FillFunctor<unsigned char>" > profiles/${T}_${n}_step_launch_summary.txt; head -1 profiles/${T}_${n}_step_launch_summary.txt; done
