# usage: bash scripts/ncu_list.sh <tag> <python args...>   -> gpurun_out/launches_<tag>.csv (per-launch device times)
tag=$1; shift
python "$@" > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python "$@" > gpurun_out/ncu_$tag.log 2>&1
echo rc=$?
