# Final ncu captures of round 2 (run under gpurun; each capture only after the same command ran clean without ncu).
# The .ncu-rep files are summarised on the box (scripts/ncu_all.py) and deleted: gpurun copies back at most 64 MiB.
set -x
O=gpurun_out
cap() {  # name, kernel regex, skip, count, script args...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 300 python "$@" > $O/plain_$name.log 2>&1 || return 1
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/$name python "$@" > $O/ncu_$name.log 2>&1
    python scripts/ncu_all.py $O/$name.ncu-rep > $O/$name.summary.txt 2>&1
    python scripts/ncu_summary.py $O/$name.ncu-rep > $O/$name.first_kernel.txt 2>&1
    rm -f $O/$name.ncu-rep
}
case "$1" in
  edge)
    cap r02b_k2 edge_fwd_ 3 3 scripts/k2_one.py
    cap r02b_k2_fused edge_fwd_staged 2 1 scripts/k2_one.py fused
    cap r02b_k2b edge_bwd_ 4 4 scripts/k2b_one.py fused
    cap r02b_narrow narrow 3 3 scripts/narrow_one.py
    cap r02b_lin_bwd lin_bwd 2 2 scripts/lin_bwd_one.py
    ;;
  k1)
    cap r02b_k1 simknn_stage1_kernel 1 1 scripts/k1_build_one.py
    cap r02b_k1_arxiv simknn_stage1_kernel 1 1 scripts/k1_build_one.py arxiv-year
    ;;
  launches)
    timeout 600 python bench.py --skip-configs --skip-cpu --steps 2 --warmup 1 > $O/bench_for_launches.json 2> $O/bench_for_launches.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02b_bench_pokec.csv python bench.py --skip-configs --skip-cpu --steps 2 --warmup 1 > $O/ncu_e.log 2>&1
    ;;
esac
ls -la $O | tail -20
