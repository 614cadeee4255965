# ncu --set full capture of one kernel, exported as the raw-page CSV (the .ncu-rep
# embeds the whole fat binary and is too large to bring back)
# usage: capture_raw.sh <name> <kernel regex> <command...>
name=$1; shift; pat=$1; shift
"$@" > gpurun_out/${name}_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:$pat -s 1 -c 1 -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
rm -f /tmp/$name.ncu-rep
