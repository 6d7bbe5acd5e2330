import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        ks={k['kernel']:round(k['ms_per_step'],1) for k in d['kernels'][:6]}
        print(sys.argv[1], round(d['ms_per_step'],1), ks, d['workload_stats'].get('spec_windows'))
