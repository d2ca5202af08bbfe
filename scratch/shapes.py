import json,sys,subprocess,os
os.environ["LCGAN_PROFILE_SHAPES"]="1"
sys.argv=["bench.py","--steps","8","--warmup","3","--no-cpu-baseline","--no-e2e"]+sys.argv[1:]
sys.path.insert(0,'.')
import bench
orig=bench.roofline_pass
def rp(step_fn,it0,_lib,pk):
    import torch
    _lib.profile_begin()
    for i in range(8): step_fn(it0+i)
    torch.cuda.synchronize()
    st=_lib.profile_end()
    rows=sorted(st.items(), key=lambda kv:-kv[1]["ms"])
    tot=sum(v["ms"] for _,v in rows)
    print("TOTAL ms/8it",tot, file=sys.stderr)
    for k,v in rows[:45]:
        tf=v["flops"]/(v["ms"]/1e3)/1e12 if v["flops"] else 0
        gb=v["bytes"]/(v["ms"]/1e3)/1e9 if v["bytes"] else 0
        print(f"{v['ms']:9.2f}ms {100*v['ms']/tot:5.1f}% n={v['n']:5d} {tf:8.1f}TF {gb:8.0f}GB/s  {k}", file=sys.stderr)
    return orig.__wrapped__(step_fn,it0,_lib,pk) if hasattr(orig,'__wrapped__') else (None,None)
bench.roofline_pass=rp
bench.main()
