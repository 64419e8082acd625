import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np, torch
    from oracle import oracle as O
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import EstimateAdvantage
    from tests.util import make_batch, assert_within_tolerance
    T, B, lam = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    b = make_batch(np.random.default_rng(0), T, B, 0.05)
    t = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
    tree = ArrayDict({"reward": t["reward"], "done": t["done"], "agent_info": {"value": t["value"]},
                      "bootstrap_value": t["bootstrap_value"], "advantage": torch.zeros_like(t["value"]),
                      "return_": torch.zeros_like(t["value"])})
    EstimateAdvantage(0.99, lam, algo="chunked")(tree)
    torch.cuda.synchronize()
    adv, ret = np.empty_like(b["value"]), np.empty_like(b["value"])
    (O.compute_discount_return if lam == 1.0 else O.compute_gae_advantage)(b["reward"], b["value"], b["done"], b["bootstrap_value"], 0.99, lam, adv, ret)
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv)
    assert_within_tolerance(tree["return_"].cpu().numpy(), ret)
    print("ok")
else:
    for shape in [(64, 32), (128, 32), (63, 32), (64, 16), (3, 16), (64, 96), (65, 80), (200, 1008), (129, 4096), (1024, 4096)]:
        r = subprocess.run([sys.executable, __file__, str(shape[0]), str(shape[1]), "0.95"], capture_output=True, text=True,
                           env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
        print(shape, r.stdout.strip()[-60:], r.stderr.strip()[-150:].replace("\n", " | "))
