# Floating-base measurements (one B200; run under gpurun from the repo root, also called by tools/measure_round.sh).
set -u
O=gpurun_out
# one ncu --set full capture of the first launch matching $1, summarised on the box into $O/summ_$2.txt (the .ncu-rep
# files together exceed what gpurun copies back)
cap() {
  local k=$1 stem=$2; shift 2
  ncu --set full --clock-control none --import-source on -k $k -c 1 -o $O/$stem "$@"
  python tools/ncu_summary.py $O/$stem.ncu-rep $O/summ_$stem.txt > /dev/null 2>&1
  python tools/ncu_lines.py $O/$stem.ncu-rep 2>/dev/null | awk '{ if ($4+0 >= 0.8 || $6+0 >= 0.8) print }' | cut -c1-260 >> $O/summ_$stem.txt
  rm -f $O/$stem.ncu-rep
}
python tools/sanitize.py > $O/r02_sanitize.json 2> $O/r02_sanitize.err; tail -c 300 $O/r02_sanitize.json; tail -3 $O/r02_sanitize.err
: > $O/r02_fb_bench.jsonl
for r in iiwa14_fb hyq_fb atlas_fb; do for op in rnea rnea_grad minv; do for dt in f64 f32; do
  python bench.py --robot $r --op $op --dtype $dt --batch 262144 --steps 200 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 >> $O/r02_fb_bench.jsonl
done; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r02_fb_bench.jsonl'):
    d=json.loads(l); c=d['config']; r=d['roofline']
    print('%-10s %-9s %s %.3e evals/s %.3f ms fma %.3f clk %s' % (c['robot'],c['op'],d['dtype'],d['value'],d['ms_per_step'],r['frac'] or 0,d['clocks']['samples']))
PY
: > $O/r02_fb_passes.jsonl
for r in iiwa14_fb hyq_fb atlas_fb; do python tools/bench_passes.py --reps 5 --robot $r --batch 65536 >> $O/r02_fb_passes.jsonl 2>/dev/null; done
cap regex:rnea_grad_coop r02_prof_fb_grad_coop_hyq_f64 python bench.py --robot hyq_fb --op rnea_grad --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_ncu6.log 2>&1
cap regex:minv_coop r02_prof_fb_minv_coop_hyq_f64 python bench.py --robot hyq_fb --op minv --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_ncu7.log 2>&1
