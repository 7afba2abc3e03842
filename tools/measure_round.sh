# Round-2 measurement script (one B200; run under gpurun from the repo root).  Writes gpurun_out/r02_*.
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
python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; tail -1 $O/r02_bench_default.json | cut -c1-200
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_ref.json 2>/dev/null; tail -1 $O/r02_bench_ref.json | cut -c1-160
: > $O/r02_matrix.jsonl
for r in iiwa14 hyq atlas; do for op in rnea_grad minv rnea crba fd fd_grad; do for dt in f64 f32; do
  b=1048576; if [ $r = atlas ]; then b=262144; fi
  python bench.py --robot $r --op $op --dtype $dt --batch $b --steps 100 --warmup 3 --no-cpu-baseline --no-e2e --no-quadrants 2>/dev/null | tail -1 >> $O/r02_matrix.jsonl
done; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r02_matrix.jsonl'):
    d=json.loads(l); c=d['config']; r=d['roofline']; h=d['roofline_hbm']
    print('%-7s %-9s %s B=%-8d %.3e evals/s  %.3f ms  fma %.3f  hbm %.3f clk %s' % (c['robot'],c['op'],d['dtype'],c['batch_per_gpu'],d['value'],d['ms_per_step'],r['frac'] or 0,h['frac'],d['clocks']['samples']))
PY
python tools/sweep.py --ops rnea_grad,minv --robots iiwa14,atlas > $O/r02_sweep_f64.jsonl 2> $O/r02_sweep.err
: > $O/r02_passes.jsonl
python tools/bench_passes.py --reps 5 --batch 1048576 >> $O/r02_passes.jsonl 2>/dev/null
python tools/bench_passes.py --reps 5 --robot hyq --batch 262144 >> $O/r02_passes.jsonl 2>/dev/null
python tools/bench_passes.py --reps 5 --robot atlas --batch 65536 >> $O/r02_passes.jsonl 2>/dev/null
: > $O/r02_ee_bench.jsonl
for r in iiwa14 hyq atlas; do for d in f64 f32; do
  b=1048576; if [ $r = atlas ]; then b=262144; fi
  python bench.py --op ee_grad --robot $r --dtype $d --batch $b --no-cpu-baseline --steps 20 2>/dev/null | tail -1 >> $O/r02_ee_bench.jsonl
done; done
# floating base (bench rows, canary / determinism checker, ncu captures of the two HyQ + base kernels)
bash tools/measure_fb.sh
# launch list of the default bench command (cold, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench_default.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_ncu_launch.log 2>&1
# full captures of the dominant kernels (each after its plain command above has exited 0)
cap regex:rnea_grad_chain r02_prof_chain_iiwa14_f64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-quadrants > $O/r02_ncu1.log 2>&1
cap regex:minv_tile r02_prof_minv_tile_atlas_f64 python bench.py --op minv --robot atlas --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_ncu2.log 2>&1
cap regex:minv_lane r02_prof_minv_lane_iiwa14_f64 python bench.py --op minv --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-quadrants > $O/r02_ncu3.log 2>&1
cap regex:rnea_grad_coop r02_prof_grad_coop_atlas_f64 python bench.py --robot atlas --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_ncu4.log 2>&1
cap regex:grad_fpass_level r02_prof_grad_fpass_level_iiwa14_f64 python tools/bench_passes.py --reps 1 --batch 1048576 > $O/r02_ncu5.log 2>&1
cap regex:grad_fpass_level r02_prof_grad_fpass_level_atlas_f64 python tools/bench_passes.py --reps 1 --robot atlas --batch 65536 > $O/r02_ncu5b.log 2>&1
ls -la $O/summ_*.txt
