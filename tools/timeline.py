"""Device timeline of a callable (torch.profiler / CUPTI): span, busy and idle time of the device, the largest gaps with
the activities either side, and the time per kernel name.  Used by tools/gaps_fourier.py and `bench.py --timeline FILE`."""
import torch
from torch.profiler import ProfilerActivity, profile


def device_timeline(fn, title="", top_gaps=20, top_kernels=16, min_gap_us=20.0):
    """Runs fn() once under the profiler (fn must have been warmed up) and returns the report as text."""
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if not ev:
        return f"{title}: no device activity recorded"
    ev.sort(key=lambda e: e.time_range.start)
    # union of the busy intervals (activities on several streams may overlap)
    busy, cur_s, cur_e = 0.0, ev[0].time_range.start, ev[0].time_range.end
    gaps = []
    prev = ev[0]
    for e in ev[1:]:
        s, t = e.time_range.start, e.time_range.end
        if s > cur_e:
            if s - cur_e > min_gap_us:
                gaps.append((s - cur_e, prev.name[:70], e.name[:70]))
            busy += cur_e - cur_s
            cur_s, cur_e = s, t
        else:
            cur_e = max(cur_e, t)
        if t >= cur_e:
            prev = e
    busy += cur_e - cur_s
    span = max(e.time_range.end for e in ev) - ev[0].time_range.start
    out = [f"{title}: {len(ev)} device activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, "
           f"idle {(span - busy) / 1e3:.2f} ms ({100.0 * (span - busy) / span:.1f} %)"]
    for g, a, b in sorted(gaps, reverse=True)[:top_gaps]:
        out.append(f"{g / 1e3:9.3f} ms idle  after {a}  before {b}")
    by = {}
    for e in ev:
        k = e.name[:70]
        c = by.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += e.time_range.end - e.time_range.start
    for k, (cnt, v) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top_kernels]:
        out.append(f"{v / 1e3:9.2f} ms {cnt:5d} x  {k}")
    return "\n".join(out)
