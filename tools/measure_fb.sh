# Floating-base measurements (one B200; run under gpurun from the repo root, also called by tools/measure_round.sh).
set -u
O=gpurun_out
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
ncu --set full --clock-control none --import-source on -k regex:rnea_grad_coop -c 1 -o $O/r02_prof_fb_grad_coop_hyq_f64 python bench.py --robot hyq_fb --op rnea_grad --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:minv_coop -c 1 -o $O/r02_prof_fb_minv_coop_hyq_f64 python bench.py --robot hyq_fb --op minv --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_ncu7.log 2>&1
ls -la $O/*fb*.ncu-rep
