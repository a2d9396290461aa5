python tools/prof_cmd.py big otf sheppard > gpurun_out/plain_lines.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lines.csv python tools/prof_cmd.py big otf sheppard > gpurun_out/ncu_lines.log 2>&1
echo rc=$?
