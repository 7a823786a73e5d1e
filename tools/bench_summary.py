import json, sys
for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(fn, "unreadable", e); continue
    print(f"{fn}: value {d['value']:.0f} fps ({d['ms_per_step']:.1f} ms/step)  e2e {d['e2e']['value']:.0f} fps  launches {d['gpu_launches']}  B={d['config']['frames_per_step']}")
    r = d['roofline']
    print(f"   roofline: {r['achieved']:.1f} GB/s = {r['frac']*100:.2f}% of {r['peak']}  ({r['ms_per_launch']:.2f} ms/launch)")
    print("   stage ms/step", r['stage_ms_per_step'])
    print("   clocks", d['clocks'])
    if d.get('cpu_baseline'):
        c = d['cpu_baseline']; print(f"   cpu {c['value']:.0f} fps on {c['cores']} cores; 1 thread {c['single_thread_frames_per_s']:.1f}")
