# ncu --set full capture of ONE launch of one kernel: raw + source pages exported as CSV on the box
# (small, always brought back) and the .ncu-rep itself when it is below 40 MiB.
# usage: capture_rep.sh <name> <kernel regex> <launch-skip> <command...>
name=$1; shift; pat=$1; shift; skip=$1; shift
ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
sz=$(stat -c %s /tmp/$name.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 0 ]; then cp /tmp/$name.ncu-rep gpurun_out/; fi
rm -f /tmp/$name.ncu-rep
