"""Would splitting a cfg2 batch into independent sub-batches on parallel streams hide the one-wave tails? (development aid)
One S0 pipeline of 16 crops vs k pipelines of 16 / k crops replayed concurrently on k streams."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import resr_b200
ip = resr_b200.imgproc
dev = torch.device("cuda")
H = W = 256


def kernels(B, seed=0):
    P = dict(resr_b200.plan.DEGRADATION_MODEL_PARAMETERS)
    P["sinc_kernel_probability3"] = 1.0
    random.seed(seed); np.random.seed(seed)
    return ip.synthesize_degradation_kernels(B, P, dev)


k1, k2, sk = kernels(16)
hr = torch.rand(16, 3, H, W, device=dev)
for parts in (1, 2, 4):
    b = 16 // parts
    pipes, streams = [], []
    for i in range(parts):
        sl = slice(i * b, (i + 1) * b)
        plan = resr_b200.plan.canonical_plan_s0(b, H, W, seed=i)
        pipes.append(ip.DegradePipeline(hr[sl], k1[sl], k2[sl], sk[sl], plan))
        streams.append(torch.cuda.Stream(device=dev))
    cur = torch.cuda.current_stream(dev)

    def run():
        for p, s in zip(pipes, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                p()
        for s in streams:
            cur.wait_stream(s)

    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    print(f"S0 16x3x256x256 as {parts} concurrent pipeline(s) of {b}: {e0.elapsed_time(e1)/50*1e3:.1f} us per 16 crops")
