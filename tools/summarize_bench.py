import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d = json.loads(line)
        e = d.get('e2e')
        print('%-48s | %.3e evals/s | %.3f ms | fp frac %.3f | hbm frac %.3f | e2e %s | clk %s' % (
            d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'] or 0,
            d['roofline_hbm']['frac'], e and '%.3e' % e['value'], d.get('clocks') and d['clocks'].get('sm_mhz')))
    elif not line.startswith('{'):
        print(line.rstrip()[:220])
