"""A few README-sized (32-row) fit steps of the c1 flow, for an ncu launch list (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probaforms_b200.models import RealNVP
from sklearn.datasets import make_moons
Xm, ym = make_moons(n_samples=1000, noise=0.1, random_state=0)
torch.manual_seed(0)
m = RealNVP(lr=0.01, n_epochs=2)
m.fit(Xm, ym.reshape(-1, 1))
torch.cuda.synchronize()
print("ok", float(torch.stack(m.loss_history)[-32:].mean()))
