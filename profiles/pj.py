import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    if 'roofline' in d: print(round(d["value"]), 'qps', round(d["roofline"]["kernel_ms"],2), 'ms frac', round(d["roofline"]["frac"],4), 'exp/q', round(d["roofline"]["expansions_per_query"]), 'e2e', round(d['e2e']['value']), 'K2 stream', d.get('fastscan_stream', {}).get('achieved'), d.get('fastscan_stream', {}).get('frac'))
    else: print(d.get('impl'), d.get('value'))
