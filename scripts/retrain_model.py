#!/usr/bin/env python
"""
Train the aslnn surrogate and write its weights in the reference's .npy layout (weights%i.npy / biases%i.npy),
the counterpart of /root/reference/scripts/retrain_model.py:30-33 (the reference does not ship its
`trained_data/`).  Needs a GPU: the training curves come from the aslrest CUDA kernel.

    python scripts/retrain_model.py [--out trained_data] [--examples 200000] [--steps 6000]

Optimiser: full-batch Adam with a decaying step instead of the reference's 7e6 plain-SGD mini-batch steps
(same network, same loss, same training distribution: t~U(1,5), delttiss~U(0.1,3), ftiss=1; aslnn.py:191-199).
`AslNNModel(train_save=...)` still offers the reference's SGD trainer.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svb import DataModel                      # noqa: E402
from svb_models_asl import AslNNModel          # noqa: E402

TIS = [2.05, 2.3, 2.55, 2.8, 3.05, 3.3]
OPTIONS = {"tau": 1.8, "t1b": 1.6, "casl": True, "repeats": 1, "t1": 1.3}     # retrain_model.py:10-24


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "trained_data"))
    ap.add_argument("--examples", type=int, default=200000)
    ap.add_argument("--steps", type=int, default=6000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    dm = DataModel(np.zeros((1, 6), dtype=np.float32))
    model = AslNNModel(dm, tis=TIS, train_delttiss_max=3.0, **OPTIONS)
    x_train, x_test, y_train, y_test = model._get_training_data(a.examples, seed=a.seed)
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(a.seed)
    layers = [(2, 10), (10, 10), (10, 1)]
    ws = [(torch.randn(i, o, generator=gen) / np.sqrt(i)).to(dev).requires_grad_(True) for i, o in layers]
    bs = [torch.zeros(1, o, device=dev, requires_grad=True) for _, o in layers]
    x, y = torch.as_tensor(x_train, device=dev), torch.as_tensor(y_train, device=dev).reshape(-1, 1)
    opt = torch.optim.Adam(ws + bs, lr=0.01)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, a.steps, eta_min=1e-4)
    for step in range(a.steps):
        h = torch.tanh(torch.tanh(x @ ws[0] + bs[0]) @ ws[1] + bs[1]) @ ws[2] + bs[2]
        loss = torch.mean((y - h) ** 2)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if step % 500 == 0 or step == a.steps - 1:
            print("step %5d  mse %.3e" % (step, float(loss)), flush=True)
    model.trained_weights = [w.detach().cpu().numpy() for w in ws]
    model.trained_biases = [b.detach().cpu().numpy() for b in bs]
    pred = model._ievaluate_nn(x_test)                      # through the CUDA evaluate kernel
    r2 = 1.0 - float(np.sum((y_test - pred) ** 2)) / float(np.sum((y_test - y_test.mean()) ** 2))
    print("test r^2 = %.6f  max abs err = %.4f (peak signal %.3f)" % (r2, np.abs(y_test - pred).max(), y_test.max()))
    model._save_nn(a.out)
    print("saved to", a.out)


if __name__ == "__main__":
    main()
