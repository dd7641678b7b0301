"""Where does a batched env.step() spend time besides the integrator kernel?  (GPU box diagnostic)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pdecontrolgym_b200 as pcg
import bench

dev = torch.device("cuda:0")
E = 65536
def mk(**kw):
    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=E, device=dev, precision="fp64",
                   reward_class=pcg.NormReward(bench.REWARD_NT, norm="2"), **{**bench.RD_KW, **kw})
    ic = torch.empty(E, 1, dtype=torch.float64, device=dev).uniform_(1, 10) * torch.ones(1, 201, dtype=torch.float64, device=dev)
    env.reset(options={"init_condition": ic, "beta": bench.beta_fn()})
    return env
K = 200
acts = torch.empty(K + 10, E, device=dev).uniform_(-1, 1)
def loop(env, events=False, retsum=False, sampler=False):
    for i in range(10): env.step(acts[i])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ret = torch.zeros(E, dtype=torch.float64, device=dev)
    sm = None
    if sampler:
        sm = bench.ClockSampler(0); sm.start(); time.sleep(0.3)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); tw0 = time.time()
    s.record()
    for i in range(K):
        if events: env._kernel_events = ev[i]
        o, r, te, tr, _ = env.step(acts[10 + i])
        if retsum: ret += r
    env._kernel_events = None
    e.record()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    if sm: print("   clocks", sm.stop(tw0, time.time()))
    return s.elapsed_time(e) / K, 1e3 * t_cpu / K
for name, kw in (("default(copy,autoreset)", dict(autoreset=True)), ("bare", dict(autoreset=False, copy_obs=False))):
    for opts in (dict(), dict(events=True), dict(retsum=True), dict(sampler=True), dict(events=True, retsum=True, sampler=True)):
        env = mk(**kw)
        g, c = loop(env, **opts)
        print(f"{name:26s} {str(opts):60s} gpu {g:7.3f} ms/step  cpu-issue {c:7.3f} ms/step", flush=True)
        del env
