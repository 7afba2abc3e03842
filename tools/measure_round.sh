python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-300
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; tail -1 gpurun_out/bench_ref.json | cut -c1-200
: > gpurun_out/matrix.jsonl
for r in iiwa14 hyq atlas; do for op in rnea_grad minv rnea crba fd fd_grad; do for dt in f64 f32; do
  b=1048576; if [ $r = atlas ]; then b=262144; fi
  python bench.py --robot $r --op $op --dtype $dt --batch $b --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 >> gpurun_out/matrix.jsonl
done; done; done
python - <<'PY'
import json
for l in open('gpurun_out/matrix.jsonl'):
    d=json.loads(l); c=d['config']; r=d['roofline']; h=d['roofline_hbm']
    print('%-7s %-9s %s B=%-8d %.3e evals/s  %.3f ms  fma %.3f  hbm %.3f' % (c['robot'],c['op'],d['dtype'],c['batch_per_gpu'],d['value'],d['ms_per_step'],r['frac'] or 0,h['frac']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rnea_grad_coop -c 1 -o gpurun_out/prof_grad_coop_iiwa3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu12.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:minv_lane -c 1 -o gpurun_out/prof_minv_lane_iiwa2 python bench.py --op minv --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:minv_hybrid -c 1 -o gpurun_out/prof_minv_hybrid_atlas4 python bench.py --op minv --robot atlas --batch 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu14.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fd_apply_mma -c 1 -o gpurun_out/prof_fd_mma_atlas2 python bench.py --op fd_grad --robot atlas --batch 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu19.log 2>&1
python tools/bench_passes.py --reps 5 --batch 1048576 > gpurun_out/passes.jsonl 2>/dev/null
python tools/bench_passes.py --reps 5 --robot hyq --batch 262144 >> gpurun_out/passes.jsonl 2>/dev/null
python tools/bench_passes.py --reps 5 --robot atlas --batch 65536 >> gpurun_out/passes.jsonl 2>/dev/null
ls -la gpurun_out/*.ncu-rep | tail -5

# ---- rows SURVEY.md 8f marks "next": end-effector kinematics and the floating base ------------------
: > gpurun_out/ee_bench.jsonl; : > gpurun_out/fb_bench.jsonl
for r in iiwa14 hyq atlas; do for d in f64 f32; do
  b=1048576; if [ $r = atlas ]; then b=262144; fi
  python bench.py --op ee_grad --robot $r --dtype $d --batch $b --no-cpu-baseline --steps 20 2>/dev/null | tail -1 >> gpurun_out/ee_bench.jsonl
done; done
for r in iiwa14_fb hyq_fb atlas_fb; do for op in rnea rnea_grad minv; do for d in f64 f32; do
  python bench.py --robot $r --op $op --dtype $d --batch 262144 --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/fb_bench.jsonl
done; done; done
ncu --set full --clock-control none --import-source on -k regex:ee_pose_kernel -c 1 -o gpurun_out/prof_ee_iiwa2 python bench.py --op ee_grad --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ee.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fb_rnea_grad -c 1 -o gpurun_out/prof_fb_grad_hyq2 python bench.py --robot hyq_fb --op rnea_grad --batch 262144 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fb_minv -c 1 -o gpurun_out/prof_fb_minv_hyq python bench.py --robot hyq_fb --op minv --batch 262144 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fb2.log 2>&1
