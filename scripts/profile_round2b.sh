# Final ncu captures of round 2 (run under gpurun; each capture only after the same command ran clean without ncu).
set -x
O=gpurun_out
timeout 200 python scripts/k2_one.py > $O/plain_k2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:edge_fwd_ -s 3 -c 3 -o $O/r02b_k2 python scripts/k2_one.py > $O/ncu_a.log 2>&1
timeout 200 python scripts/k2_one.py fused > $O/plain_k2f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:edge_fwd_staged -s 2 -c 1 -o $O/r02b_k2_fused python scripts/k2_one.py fused > $O/ncu_b.log 2>&1
timeout 200 python scripts/k2b_one.py fused > $O/plain_k2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:edge_bwd_ -s 4 -c 4 -o $O/r02b_k2b python scripts/k2b_one.py fused > $O/ncu_c.log 2>&1
timeout 200 python scripts/narrow_one.py > $O/plain_narrow.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:narrow -s 3 -c 3 -o $O/r02b_narrow python scripts/narrow_one.py > $O/ncu_n.log 2>&1
timeout 200 python scripts/lin_bwd_one.py > $O/plain_linbwd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lin_bwd -s 2 -c 2 -o $O/r02b_lin_bwd python scripts/lin_bwd_one.py > $O/ncu_l.log 2>&1
timeout 200 python scripts/k1_build_one.py > $O/plain_k1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simknn_stage1_kernel -s 1 -c 1 -o $O/r02b_k1 python scripts/k1_build_one.py > $O/ncu_d.log 2>&1
timeout 200 python scripts/k1_build_one.py arxiv-year > $O/plain_k1a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simknn_stage1_kernel -s 1 -c 1 -o $O/r02b_k1_arxiv python scripts/k1_build_one.py arxiv-year > $O/ncu_da.log 2>&1
timeout 600 python bench.py --skip-configs --skip-cpu --steps 2 --warmup 1 > $O/bench_for_launches.json 2> $O/bench_for_launches.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02b_bench_pokec.csv python bench.py --skip-configs --skip-cpu --steps 2 --warmup 1 > $O/ncu_e.log 2>&1
ls -la $O/r02b_*.ncu-rep | tail -8
