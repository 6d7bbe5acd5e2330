import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l)
        print({k["kernel"]:round(k["ms_per_step"],1) for k in d["kernels"] if "mer" in k["kernel"] or "dup" in k["kernel"] or "park" in k["kernel"]})
        print('kernel_ms',round(d['kernel_ms_per_step'],1),'host_ms',round(d['host_ms_per_step'],1), d['stage_wall_ms'])
