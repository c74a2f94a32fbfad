"""A few c1 (README moons) fit steps and sample calls for an ncu launch list (development aid)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVP
rng = np.random.default_rng(0)
X = rng.normal(size=(320, 2)); C = (rng.random((320, 1)) > 0.5).astype(np.float64)
torch.manual_seed(0)
m = RealNVP(lr=0.01, n_epochs=2)
m.fit(X, C)
m.sample(C)
torch.cuda.synchronize()
