import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    if 'roofline' in d:
        r=d["roofline"]
        print(round(d["value"]), 'qps', round(r["kernel_ms"],2), 'ms frac', round(r["frac"],4), '| one at a time:', round(d.get("value_one_batch_at_a_time",0)), 'qps',
              round(r.get("kernel_ms_one_at_a_time",0),2), 'ms frac', round(r.get("frac_one_at_a_time",0),4), '| exp/q', round(r["expansions_per_query"]),
              '| e2e', round(d['e2e']['value']), 'sync', round(d['e2e'].get('value_one_batch_at_a_time',0)),
              '| K2', d.get('fastscan_stream', {}).get('achieved'), d.get('fastscan_stream', {}).get('frac'))
    else: print(d.get('impl'), d.get('value'))
