"""
``aslnn`` plugin: neural-network surrogate of the ASL kinetic curve.

Host-side mirror of ``AslNNModel`` (/root/reference/svb_models_asl/aslnn.py): same options (:38-59), the
two parameters ``ftiss`` (LogNormal) and ``delttiss`` (FoldedNormal) (:73-81), ``signal = ftiss * MLP([t, delttiss])``
with MLP = 2 -> 10 tanh -> 10 tanh -> 1 (:93-126, :229-260), weights in the reference's ``weights%i.npy`` /
``biases%i.npy`` layout (:211-227, :326-340), trainer on AslRestModel-simulated curves (:172-209, :262-299).
The forward pass and its delttiss-derivative run inside the fused CUDA kernel (csrc/model_nn.h).
"""
import math
import os

import numpy as np

from ..svbcompat.model import Model, ModelOption
from ..svbcompat.parameter import get_parameter
from ..svbcompat.utils import NP_DTYPE, ValueList
from .. import _lib as L
from .aslrest import AslRestModel, __version__

LAYERS = [(2, 10), (10, 10), (10, 1)]           # aslnn.py:238-240


# fused step on tcgen05 by default: 1.33 G against 1.15 G voxel-iters/s for the FP32-pipe kernel, same results to
# float32 rounding (profiles/r2_notes.md section 4)
TENSOR_CORE_STEP_DEFAULT = True


class AslNNModel(Model):
    """ASL resting state model using NN for evaluation"""

    OPTIONS = [
        ModelOption("tau", "Bolus duration", units="s", clargs=("--tau", "--bolus"), type=float, default=1.8),
        ModelOption("casl", "Data is CASL/pCASL", type=bool, default=False),
        ModelOption("att", "Bolus arrival time", units="s", type=float, default=1.3),
        ModelOption("attsd", "Bolus arrival time prior std.dev.", units="s", type=float, default=None),
        ModelOption("t1", "Tissue T1 value", units="s", type=float, default=1.3),
        ModelOption("t1b", "Blood T1 value", units="s", type=float, default=1.65),
        ModelOption("tis", "Inversion times", units="s", type=ValueList(float)),
        ModelOption("plds", "Post-labelling delays (for CASL instead of TIs)", units="s", type=ValueList(float)),
        ModelOption("repeats", "Number of repeats - single value or one per TI/PLD", units="s", type=ValueList(int),
                    default=1),
        ModelOption("slicedt", "Increase in TI/PLD per slice", units="s", type=float, default=0),
        ModelOption("pc", "Blood/tissue partition coefficient", type=float, default=0.9),
        ModelOption("fcalib", "Perfusion value to use in estimation of effective T1", type=float, default=0.01),
        ModelOption("train_ti_max", "Maximum TI to train for", type=float, default=20.0),
        ModelOption("train_delttiss_max", "Maximum value of ATT to train for", type=float, default=3.0),
        ModelOption("train_lr", "Training learning rate", type=float, default=0.001),
        ModelOption("train_steps", "Training steps", type=int, default=30000),
        ModelOption("train_batch_size", "Training batch size", type=int, default=100),
        ModelOption("train_examples", "Number of training examples", type=int, default=500),
        ModelOption("train_save", "Directory to save trained model weights to"),
        ModelOption("train_load", "Directory to load trained model weights from"),
        # extensions of this engine (no counterpart in the reference): where the 10x10 layer runs
        ModelOption("use_tensor_cores", "evaluate(): 10x10 layer on the tensor cores (csrc/nn_tc.cu)", type=bool, default=False),
        ModelOption("tensor_core_step", "fused SVB step: the two 10x10 products per row on the tensor cores "
                    "(csrc/model_nn_tc.cuh)", type=bool, default=TENSOR_CORE_STEP_DEFAULT),
    ]

    KIND = L.MODEL_ASLNN

    def __init__(self, data_model, **options):
        Model.__init__(self, data_model, **options)
        if self.plds is not None:
            self.tis = [self.tau + pld for pld in self.plds]
        if self.tis is None:
            raise ValueError("Either TIs or PLDs must be given")
        if self.attsd is None:
            self.attsd = 1.0 if len(self.tis) > 1 else 0.1
        reps = self.repeats
        if isinstance(reps, (int, np.integer)):
            reps = [int(reps)]
        self.repeats = int(list(reps)[0])                                # aslnn.py:68-71
        self.params = [
            get_parameter("ftiss", dist="LogNormal", mean=1.5, prior_var=1e6, post_var=1.5,
                          post_init=self._init_flow, **options),
            get_parameter("delttiss", dist="FoldedNormal", mean=self.att, var=self.attsd ** 2, **options),
        ]
        self.trained_weights = None
        self.trained_biases = None
        self.train_r2 = None
        if self.train_save:
            self._init_nn()

    def __str__(self):
        return "ASL neural network model: %s" % __version__

    # ---- Model API ----
    def evaluate(self, params, tpts):
        """params: ftiss, delttiss each [M,S,1] (or one [2,M,S,1]); tpts [1,1,N] or [M,1,N] -> [M,S,N] CUDA tensor"""
        from ..ops import evaluate_model, nn_evaluate_tc
        if self.trained_weights is None:
            self._init_nn()
        # Two kernels give the same result: the FP32-pipe one (csrc/model_nn.h) and the stand-alone tcgen05 one
        # (csrc/nn_tc.cu, A tiles through shared memory).  For the forward pass alone - one 10x10 product per row,
        # nothing else to amortise the tile round trip over - the FP32 pipe is faster (65-69 against 49-52 G rows/s,
        # profiles/r2_notes.md section 4) and is the default; `use_tensor_cores=True` selects the other.  The fused
        # SVB step is a different matter: see `tensor_core_step` / kernel_model().
        if getattr(self, "use_tensor_cores", False):
            return nn_evaluate_tc(self, params, tpts)
        return evaluate_model(self, params, tpts)

    def tpts(self):
        n_expected = len(self.tis) * self.repeats
        if self.data_model.n_tpts != n_expected:
            raise ValueError("ASL model configured with %i time points, but data has %i"
                             % (n_expected, self.data_model.n_tpts))
        base = np.repeat(np.asarray(self.tis, dtype=np.float64), self.repeats)
        if self.slicedt > 0:
            # mask the per-voxel timings like aslrest does (the reference forgets to)
            z = self.data_model.voxel_coords()[:, 2].astype(np.float64)
            return (base[None, :] + (z * self.slicedt)[:, None]).astype(NP_DTYPE)
        return base.reshape(1, -1).astype(NP_DTYPE)

    def kernel_model(self, device_array=None):
        if self.trained_weights is None:
            self._init_nn()
        m = L.Model()
        m.kind = self.KIND
        m.flags = L.F_CASL if self.casl else 0
        if getattr(self, "tensor_core_step", False):
            m.flags |= L.F_NN_TC             # fused step: the two 10x10 products on tcgen05 (csrc/model_nn_tc.cuh)
        m.tau, m.t1b = self.tau, self.t1b
        packed = self.packed_weights()
        m.nn_weights = packed.ctypes.data                                # HOST pointer, copied into kernel args
        return m, [packed]

    def packed_weights(self):
        """W0[2][10] b0[10] W1[10][10] b1[10] W2[10] b2[1] as one float32 vector (include/svbasl.h)."""
        parts = []
        for w, b in zip(self.trained_weights, self.trained_biases):
            parts += [np.asarray(w, dtype=np.float32).reshape(-1), np.asarray(b, dtype=np.float32).reshape(-1)]
        packed = np.ascontiguousarray(np.concatenate(parts))
        if packed.size != 151:
            raise ValueError("aslnn expects a 2-10-10-1 network (151 weights), got %i" % packed.size)
        return packed

    def _init_flow(self, _param, _t, data):
        # aslnn.py:143-147 takes the plain mean; ftiss is LogNormal here, so a non-positive mean would make the
        # initial log-mean NaN - floored like aslrest's initialiser (aslrest.py:467)
        st = getattr(data, "device_stats", None)               # reduced on the GPU by the engine (ops.InitData)
        mean = np.asarray(st["mean_t"]) if st is not None else np.asarray(data).mean(axis=1)
        return np.maximum(mean, 0.1).astype(NP_DTYPE), None

    # ---- weights: load / save / train ----
    def _init_nn(self):
        self.log.info("Initializing neural-network based ASL model")
        if self.train_load:
            self._load_nn(self.train_load)
            return
        x_train, x_test, y_train, y_test = self._get_training_data(self.train_examples)
        self._train_nn(x_train, y_train, self.train_steps, self.train_lr, self.train_batch_size)
        pred = self._ievaluate_nn(x_test)
        ss_res = float(np.sum((y_test - pred) ** 2))
        ss_tot = float(np.sum((y_test - y_test.mean()) ** 2))
        self.train_r2 = 1.0 - ss_res / ss_tot
        self.log.info(" - Trained model using %i steps and %.5f learning rate - accuracy %.3f",
                      self.train_steps, self.train_lr, self.train_r2)
        if self.train_save:
            self._save_nn(self.train_save)

    def _load_nn(self, load_dir):
        ws, bs = [], []
        idx = 0
        while True:
            wf = os.path.join(load_dir, "weights%i.npy" % idx)
            bf = os.path.join(load_dir, "biases%i.npy" % idx)
            if not os.path.exists(wf) and not os.path.exists(bf):
                break
            if not (os.path.exists(wf) and os.path.exists(bf)):
                raise RuntimeError("For layer %i, could not find both weights and biases" % idx)
            ws.append(np.load(wf))
            bs.append(np.load(bf))
            idx += 1
        if [tuple(w.shape) for w in ws] != LAYERS:
            raise RuntimeError("%s does not hold a 2-10-10-1 network (found %s)" % (load_dir, [w.shape for w in ws]))
        self.trained_weights, self.trained_biases = ws, bs
        self.log.info(" - Loaded %i layers from %s", len(ws), load_dir)

    def _save_nn(self, save_dir):
        if self.trained_weights is None:
            raise RuntimeError("Can't save model before it has been trained!")
        os.makedirs(save_dir, exist_ok=True)
        for idx, (w, b) in enumerate(zip(self.trained_weights, self.trained_biases)):
            np.save(os.path.join(save_dir, "weights%i.npy" % idx), np.asarray(w, dtype=np.float32))
            np.save(os.path.join(save_dir, "biases%i.npy" % idx), np.asarray(b, dtype=np.float32).reshape(1, -1))

    def _get_training_data(self, n, seed=None):
        """(t, delttiss) -> analytic signal at ftiss = 1 from AslRestModel (aslnn.py:172-209)."""
        from ..svbcompat.data import DataModel
        rng = np.random.default_rng(seed)
        dm = DataModel(np.zeros((1, len(self.tis)), dtype=np.float32))
        model = AslRestModel(dm, tis=self.tis, tau=self.tau, t1b=self.t1b, casl=True, repeats=1, t1=self.t1)
        t = rng.uniform(1.0, 5.0, size=n)
        delt = rng.uniform(0.1, self.train_delttiss_max, size=n)
        params = np.zeros((2, n, 1), dtype=np.float32)
        params[0], params[1, :, 0] = 1.0, delt
        y = model.ievaluate(params, t[:, None].astype(np.float32))[:, 0, 0]
        x = np.stack([t, delt], axis=1).astype(np.float32)
        n_test = int(round(0.3 * n))                                     # 70/30 split (aslnn.py:207)
        perm = rng.permutation(n)
        test, train = perm[:n_test], perm[n_test:]
        return x[train], x[test], y[train].astype(np.float32), y[test].astype(np.float32)

    def _train_nn(self, x_train, y_train, steps, learning_rate, batch_size=100, seed=0):
        """Plain SGD on the MSE over strided mini-batches (aslnn.py:262-299).  Offline utility: runs in
        PyTorch on the GPU when there is one (the inference path never comes here)."""
        import torch
        dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        gen = torch.Generator().manual_seed(seed)
        ws = [torch.randn(i, o, generator=gen).to(dev).requires_grad_(True) for i, o in LAYERS]   # N(0,1)
        bs = [torch.full((1, o), 0.1, device=dev, requires_grad=True) for _, o in LAYERS]         # 0.1
        x = torch.as_tensor(x_train, device=dev)
        y = torch.as_tensor(y_train, device=dev).reshape(-1, 1)
        n_batches = int(math.ceil(x.shape[0] / batch_size))
        for step in range(steps):
            total = 0.0
            for b in range(n_batches):
                xb, yb = x[b::n_batches], y[b::n_batches]
                h = torch.tanh(xb @ ws[0] + bs[0])
                h = torch.tanh(h @ ws[1] + bs[1])
                loss = torch.mean(torch.sum((yb - (h @ ws[2] + bs[2])) ** 2, dim=1))
                grads = torch.autograd.grad(loss, ws + bs)
                with torch.no_grad():
                    for p, g in zip(ws + bs, grads):
                        p -= learning_rate * g
                total += float(loss) if step % 100 == 0 else 0.0
            if step % 100 == 0:
                self.log.info(" - Step %i, cost %f", step, total / n_batches)
        self.trained_weights = [w.detach().cpu().numpy() for w in ws]
        self.trained_biases = [b.detach().cpu().numpy() for b in bs]

    def _ievaluate_nn(self, x):
        """MLP output for rows of (t, delttiss) through the CUDA evaluate kernel with ftiss = 1."""
        x = np.asarray(x, dtype=np.float32)
        n = x.shape[0]
        params = np.ones((2, n, 1), dtype=np.float32)
        params[1, :, 0] = x[:, 1]
        return self.ievaluate(params, x[:, :1])[:, 0, 0]
