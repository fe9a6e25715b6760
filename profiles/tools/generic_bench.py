import sys, time, torch
sys.path.insert(0, "/root/repo")
from cleanmarl_b200.mappo import MAPPO, Args
for kw in (dict(actor_num_layers=2, critic_hidden_dim=128), dict(n_agents=5), dict(actor_hidden_dim=128, critic_hidden_dim=128, actor_num_layers=2, critic_num_layers=2)):
    tr = MAPPO(Args(batch_size=4096, seed=1, **kw))
    for _ in range(4): tr.iteration()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): tr.iteration()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    n = 4096 * 25 * tr.engine.shapes.n_agents
    print(kw, f"{dt*1e3:.2f} ms/iteration, {n/dt:.3e} agent-env-steps/s, graph={tr.use_graph}, params={tr.engine.n_params}", flush=True)
