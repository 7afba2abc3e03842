# Copies / summarises gpurun_out/r02_* (written by tools/measure_round.sh on the GPU box) into profiles/ (tracked).
set -u
G=gpurun_out; P=profiles
summ() {  # $1 = capture stem (tools/measure_round.sh summarises on the GPU box: gpurun_out/summ_<stem>.txt), $2 = profiles name
  if [ -f $G/summ_$1.txt ]; then cp $G/summ_$1.txt $P/$2; return 0; fi
  [ -f $G/$1.ncu-rep ] || return 0
  python tools/ncu_summary.py $G/$1.ncu-rep $P/$2 > /dev/null 2>&1
  python tools/ncu_lines.py $G/$1.ncu-rep 2>/dev/null | awk '{ if ($4+0 >= 0.8 || $6+0 >= 0.8) print }' | cut -c1-260 >> $P/$2
}
summ r02_prof_chain_iiwa14_f64 r02_grad_chain_iiwa14_f64.txt
summ r02_prof_minv_tile_atlas_f64 r02_minv_tile_atlas_f64.txt
summ r02_prof_minv_lane_iiwa14_f64 r02_minv_lane_iiwa14_f64.txt
summ r02_prof_grad_coop_atlas_f64 r02_grad_coop_atlas_f64.txt
summ r02_prof_grad_fpass_level_iiwa14_f64 r02_grad_fpass_level_iiwa14_f64.txt
summ r02_prof_grad_fpass_level_atlas_f64 r02_grad_fpass_level_atlas_f64.txt
summ r02_prof_fb_grad_coop_hyq_f64 r02_fb_grad_coop_hyq_f64.txt
summ r02_prof_fb_minv_coop_hyq_f64 r02_fb_minv_coop_hyq_f64.txt
for f in r02_matrix.jsonl r02_passes.jsonl r02_fb_passes.jsonl r02_sweep_f64.jsonl r02_ee_bench.jsonl r02_fb_bench.jsonl r02_sanitize.json r02_launches_bench_default.csv; do
  [ -f $G/$f ] && cp $G/$f $P/$f
done
[ -f $G/r02_bench_default.json ] && tail -1 $G/r02_bench_default.json > $P/r02_bench_default.json
[ -f $G/r02_bench_ref.json ] && tail -1 $G/r02_bench_ref.json > $P/r02_bench_reference_arm.json
for n in 2 4 8; do [ -f $G/bench_n${n}_r02.json ] && tail -1 $G/bench_n${n}_r02.json > $P/r02_bench_n$n.json; done
[ -f $G/pcie_n8.json ] && cp $G/pcie_n8.json $P/r02_pcie_probe_n8.json
[ -f $G/pcie_n2.json ] && cp $G/pcie_n2.json $P/r02_pcie_probe_n2.json
# SASS evidence of the bulk-copy (TMA) path: opcode counts of the kernels that use it
{
  echo "cuobjdump -sass opcode counts (UBLKCP = cp.async.bulk 1-D TMA copies, SYNCS.* = mbarrier, UTMACMDFLUSH = bulk-group commit)";
  for o in rbd_launch_pass_double rbd_launch_grad_double rbd_launch_minv_double rbd_launch_fb; do
    echo "== $o.o"; cuobjdump -sass rbdreference_b200/csrc/_build/$o.o | grep -E "UBLKCP|SYNCS|UTMA" | awk '{ i=2; if ($2 ~ /^@/) i=3; print $i }' | sed 's/;//' | sort | uniq -c | sort -rn
  done
  echo "== UBLKCP instructions per kernel"
  for o in rbd_launch_pass_double rbd_launch_grad_double rbd_launch_minv_double rbd_launch_fb; do
    cuobjdump -sass rbdreference_b200/csrc/_build/$o.o | grep -E "Function|UBLKCP" | awk '/Function/ { f=$3 } /UBLKCP/ { c[f]++ } END { for (k in c) print c[k], k }' | c++filt | sed 's/(.*//' | sort -k2
  done
} > $P/r02_sass_bulk_copy.txt
ls -la $P | grep r02
