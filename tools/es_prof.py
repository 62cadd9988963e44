"""ES-NSRA step at C5 for profiling (ncu launch list): python tools/es_prof.py [steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_exploration_b200 as ppx
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
np.random.seed(0)
es = ppx.EvolutionStrategy(obs_dim=8, n_actions=2, hidden_sizes=(64, 64), population_size=10000, sigma=0.1, learning_rate=0.01,
                           decay=0.9995, novelty_param=0.5, device=dev, noise_table_size=1 << 28, noise_seed=0)
es.noise_table()
g = torch.Generator(device=dev).manual_seed(1)
archive = torch.randn(10000, 2, dtype=torch.float64, device=dev, generator=g)
queries = torch.randn(2, 2, dtype=torch.float64, device=dev, generator=g)
fit = torch.randn(10000, dtype=torch.float64, device=dev, generator=g)
es.use_graphs = False if os.environ.get("ES_EAGER") == "1" else True
for _ in range(steps):
    if os.environ.get("ES_EAGER") == "1":
        pop = es._get_population()
        w = es.perturb_all(pop)
        _, nov = es.novelty_batch(archive, queries)
        es._update_weights(fit, pop, novelty=nov[0:1])
    else:
        es.ask(archive, queries); es.tell(fit)
torch.cuda.synchronize()
print("ok")
