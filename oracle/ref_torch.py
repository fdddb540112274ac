"""CPU oracle, level 2: torch-CPU restatement of the *exact reference graph* of 3D-R2N2's
recurrent reconstruction path (per-view encoder inside the time loop, dense convolutions
on explicitly zero-stuffed Unpool3D tensors, true-convolution kernels, no algebraic
shortcuts).  Runs in float32 (the reference's floatX, .theanorc:2) or float64.

TEST INFRASTRUCTURE ONLY: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import this.  The product package must never
route through it.

PARITY UNPINNED (see ref_numpy.py header): Theano is absent; this file is cross-checked
against the independent NumPy/SciPy level-1 oracle in tests/test_oracle.py and against
the structure/shape/init fixtures generated from the real reference sources.

All citations are relative to /root/reference.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ primitive layers
def conv2d_same(x, W, b):
    """lib/layers.py:304-333.  True convolution == correlation with the flipped kernel."""
    kh, kw = W.shape[2], W.shape[3]
    return F.conv2d(x, W.flip(2, 3), b, padding=((kh - 1) // 2, (kw - 1) // 2))


def conv3d_same(x, W, b):
    """lib/layers.py:424-442.  x (B,D,C,H,W); W (O,kd,C,kh,kw); conv3d2d flips all 3 axes."""
    kd, kh, kw = W.shape[1], W.shape[3], W.shape[4]
    w = W.permute(0, 2, 1, 3, 4).flip(2, 3, 4)
    y = F.conv3d(x.permute(0, 2, 1, 3, 4), w, b,
                 padding=((kd - 1) // 2, (kh - 1) // 2, (kw - 1) // 2))
    return y.permute(0, 2, 1, 3, 4)


def pool2d(x):
    """lib/layers.py:349-353: 2x2/2 max pool, pad 1 (never wins), floor."""
    return F.max_pool2d(x, 2, 2, 1)


def unpool3d(x):
    """lib/layers.py:375-385: zero-stuffing, value at the even corner."""
    B, D, C, H, W = x.shape
    out = x.new_zeros(B, 2 * D, C, 2 * H, 2 * W)
    out[:, ::2, :, ::2, ::2] = x
    return out


def leaky_relu(x, leakiness=0.01):
    """lib/layers.py:636-646."""
    return 0.5 * (1 + leakiness) * x + 0.5 * (1 - leakiness) * x.abs()


def fcconv3d(h, f, Wh, Wx, b):
    """lib/layers.py:489-505."""
    B, D, _, H, W = h.shape
    O = Wh.shape[0]
    conv = conv3d_same(h, Wh, None)
    fc = (f @ Wx).reshape(B, D, O, H, W)
    return conv + fc + b.view(1, 1, -1, 1, 1)


def softmax_with_loss3d(x, y):
    """lib/layers.py:583-601 (no max subtraction)."""
    e = torch.exp(x)
    s = e.sum(dim=2, keepdim=True)
    pred = e / s
    loss = torch.mean(torch.sum(-y * x, dim=2, keepdim=True) + torch.log(s))
    return pred, loss


# ------------------------------------------------------------------ parameters
N_CONVFILTER = [96, 128, 256, 256, 256, 256]
N_FC = 1024
N_DECONV = [128, 128, 128, 64, 32, 2]
N_GRU_VOX = 4


def param_spec(net_name):
    """Parameter creation order == order in weights.npy == net.params
    (lib/layers.py:78-79; models/res_gru_net.py:33-76,166-197; models/gru_net.py:33-61,125-137).
    Returns list of (name, shape, is_bias)."""
    c, d = N_CONVFILTER, N_DECONV
    spec = []

    def conv2(name, cout, cin, k):
        spec.append((name + '.W', (cout, cin, k, k), False))
        spec.append((name + '.b', (cout,), True))

    def conv3(name, cout, cin, k):
        spec.append((name + '.W', (cout, k, cin, k, k), False))
        spec.append((name + '.b', (cout,), True))

    def fcconv(name):
        spec.append((name + '.Wh', (d[0], 3, d[0], 3, 3), False))
        spec.append((name + '.Wx', (N_FC, N_GRU_VOX * d[0] * N_GRU_VOX * N_GRU_VOX), False))
        spec.append((name + '.b', (d[0],), True))

    if net_name == 'ResidualGRUNet' or net_name == 'ResidualLSTMNet':
        conv2('conv1a', c[0], 3, 7); conv2('conv1b', c[0], c[0], 3)
        conv2('conv2a', c[1], c[0], 3); conv2('conv2b', c[1], c[1], 3); conv2('conv2c', c[1], c[0], 1)
        conv2('conv3a', c[2], c[1], 3); conv2('conv3b', c[2], c[2], 3); conv2('conv3c', c[2], c[1], 1)
        conv2('conv4a', c[3], c[2], 3); conv2('conv4b', c[3], c[3], 3)
        conv2('conv5a', c[4], c[3], 3); conv2('conv5b', c[4], c[4], 3); conv2('conv5c', c[4], c[3], 1)
        conv2('conv6a', c[5], c[4], 3); conv2('conv6b', c[5], c[5], 3)
    elif net_name == 'GRUNet':
        cin = 3
        for i in range(6):
            conv2('conv%d' % (i + 1), c[i], cin, 7 if i == 0 else 3)
            cin = c[i]
    else:
        raise ValueError(net_name)
    spec.append(('fc7.W', (c[5] * 3 * 3, N_FC), False))
    spec.append(('fc7.b', (N_FC,), True))
    if net_name == 'ResidualLSTMNet':
        fcconv('t_x_s_forget'); fcconv('t_x_s_input'); fcconv('t_x_s_cell')
    else:
        fcconv('t_x_s_update'); fcconv('t_x_s_reset'); fcconv('t_x_rs')
    if net_name in ('ResidualGRUNet', 'ResidualLSTMNet'):
        conv3('conv7a', d[1], d[0], 3); conv3('conv7b', d[1], d[1], 3)
        conv3('conv8a', d[2], d[1], 3); conv3('conv8b', d[2], d[2], 3)
        conv3('conv9a', d[3], d[2], 3); conv3('conv9b', d[3], d[3], 3); conv3('conv9c', d[3], d[2], 1)
        conv3('conv10a', d[4], d[3], 3); conv3('conv10b', d[4], d[4], 3); conv3('conv10c', d[4], d[4], 3)
        conv3('conv11', d[5], d[4], 3)
    else:
        conv3('conv7', d[1], d[0], 3); conv3('conv8', d[2], d[1], 3); conv3('conv9', d[3], d[2], 3)
        conv3('conv10', d[4], d[3], 3); conv3('conv11', d[5], d[4], 3)
    return spec


def msra_std(shape):
    """lib/layers.py:33-64: n = (fan_in+fan_out)/2, fan_in = prod(shape[1:]),
    fan_out = prod(shape)/shape[1] (<=4-D) or prod(shape)/shape[2] (5-D); std = sqrt(2/n).
    (FCConv3DLayer's explicit fan_in/fan_out kwargs are clobbered at 35-36 for 2-D shapes.)"""
    shape = tuple(int(s) for s in shape)
    fan_in = np.prod(shape[1:])
    fan_out = np.prod(shape) / (shape[2] if len(shape) == 5 else shape[1])
    return math.sqrt(2.0 / ((fan_in + fan_out) / 2.0))


def init_params(net_name, seed=0):
    """Seeded version of lib/layers.py:18-79 (the reference itself is unseeded,
    layers.py:31): MSRA normal for weights, constant 0.1 for biases, drawn in creation
    order from ONE np.random.RandomState(seed).  Returns a list of float32 arrays in the
    reference's weights.npy order."""
    rng = np.random.RandomState(seed)
    out = []
    for _, shape, is_bias in param_spec(net_name):
        if is_bias:
            out.append(np.full(shape, 0.1, np.float32))
        else:
            out.append(np.asarray(rng.normal(0, msra_std(shape), shape), dtype=np.float32))
    return out


def synthetic_batch(T, B, seed_x=1234, seed_y=4321, n_vox=32, img=127):
    """ShapeNet-shaped synthetic inputs (lib/data_process.py:133-154)."""
    x = np.random.RandomState(seed_x).uniform(0, 1, (T, B, 3, img, img)).astype(np.float32)
    occ = np.random.RandomState(seed_y).rand(B, n_vox, n_vox, n_vox) < 0.1
    y = np.zeros((B, n_vox, 2, n_vox, n_vox), np.float32)
    y[:, :, 1] = occ
    y[:, :, 0] = ~occ
    return x, y


# ------------------------------------------------------------------ networks
class RefNet:
    """Restatement of models/gru_net.py / models/res_gru_net.py.  `params` is the list of
    arrays in weights.npy order; held as torch tensors (leaf, requires_grad optional)."""

    def __init__(self, net_name, params, dtype=torch.float32, requires_grad=False):
        self.net_name = net_name
        self.spec = param_spec(net_name)
        assert len(params) == len(self.spec)
        self.dtype = dtype
        self.P = []
        for (name, shape, _), a in zip(self.spec, params):
            t = torch.as_tensor(np.asarray(a)).to(dtype).clone()
            assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
            t.requires_grad_(requires_grad)
            self.P.append(t)
        self.by_name = {n: t for (n, _, _), t in zip(self.spec, self.P)}

    def p(self, layer):
        d = self.by_name
        if layer + '.W' in d:
            return d[layer + '.W'], d[layer + '.b']
        return d[layer + '.Wh'], d[layer + '.Wx'], d[layer + '.b']

    # ---- encoders (per view) -------------------------------------------------------
    def encoder_res(self, x):
        """models/res_gru_net.py:78-126."""
        c = lambda name, t: conv2d_same(t, *self.p(name))
        l = leaky_relu
        pool1 = pool2d(l(c('conv1b', l(c('conv1a', x)))))
        res2 = c('conv2c', pool1) + l(c('conv2b', l(c('conv2a', pool1))))
        pool2 = pool2d(res2)
        res3 = c('conv3c', pool2) + l(c('conv3b', l(c('conv3a', pool2))))
        pool3 = pool2d(res3)
        pool4 = pool2d(l(c('conv4b', l(c('conv4a', pool3)))))
        res5 = c('conv5c', pool4) + l(c('conv5b', l(c('conv5a', pool4))))
        pool5 = pool2d(res5)
        res6 = pool5 + l(c('conv6b', l(c('conv6a', pool5))))
        pool6 = pool2d(res6)
        flat = pool6.reshape(pool6.shape[0], -1)           # flatten(2): (C,H,W)-major
        W, b = self.p('fc7')
        return l(flat @ W + b)

    def encoder_gru(self, x):
        """models/gru_net.py:63-83: conv -> pool -> LeakyReLU, six times."""
        t = x
        for i in range(6):
            t = leaky_relu(pool2d(conv2d_same(t, *self.p('conv%d' % (i + 1)))))
        flat = t.reshape(t.shape[0], -1)
        W, b = self.p('fc7')
        return leaky_relu(flat @ W + b)

    # ---- recurrent units ------------------------------------------------------------
    def gru_step(self, h, f):
        """models/gru_net.py:87-110 / res_gru_net.py:130-151."""
        u = torch.sigmoid(fcconv3d(h, f, *self.p('t_x_s_update')))
        r = torch.sigmoid(fcconv3d(h, f, *self.p('t_x_s_reset')))
        c = torch.tanh(fcconv3d(r * h, f, *self.p('t_x_rs')))
        return u * h + (1.0 - u) * c, u

    def lstm_step(self, h, s, f):
        """No reference code (lib/layers.py:508-575 is a stub); built from FCConv3DLayer x3
        following imgs/lstm.png / README.md:36-38: forget/input gates,
        cell candidate, h = tanh(s), no output gate, all convs on h_{t-1}."""
        fg = torch.sigmoid(fcconv3d(h, f, *self.p('t_x_s_forget')))
        ig = torch.sigmoid(fcconv3d(h, f, *self.p('t_x_s_input')))
        g = torch.tanh(fcconv3d(h, f, *self.p('t_x_s_cell')))
        s_new = fg * s + ig * g
        return torch.tanh(s_new), s_new, ig

    # ---- decoders ----------------------------------------------------------------------
    def decoder_res(self, h):
        """models/res_gru_net.py:165-197."""
        c = lambda name, t: conv3d_same(t, *self.p(name))
        l = leaky_relu
        unpool7 = unpool3d(h)
        res7 = unpool7 + l(c('conv7b', l(c('conv7a', unpool7))))
        unpool8 = unpool3d(res7)
        res8 = unpool8 + l(c('conv8b', l(c('conv8a', unpool8))))
        unpool9 = unpool3d(res8)
        rect9 = l(c('conv9b', l(c('conv9a', unpool9))))
        res9 = c('conv9c', unpool9) + rect9
        rect10a = l(c('conv10a', res9))
        rect10 = l(c('conv10b', rect10a))
        res10 = c('conv10c', rect10a) + rect10
        return c('conv11', res10)

    def decoder_gru(self, h):
        """models/gru_net.py:124-137."""
        c = lambda name, t: conv3d_same(t, *self.p(name))
        l = leaky_relu
        t = l(c('conv7', unpool3d(h)))
        t = l(c('conv8', unpool3d(t)))
        t = l(c('conv9', unpool3d(t)))
        t = l(c('conv10', t))
        return c('conv11', t)

    # ---- whole net ------------------------------------------------------------------------
    def forward(self, x, y=None):
        """x: (T,B,3,127,127), y: (B,32,2,32,32) or None (zeros fed, lib/solver.py:222-226).
        Returns dict(pred, loss, update_gates (T,B,4,128,4,4), logits)."""
        x = torch.as_tensor(x).to(self.dtype)
        T, B = x.shape[0], x.shape[1]
        h = x.new_zeros(B, N_GRU_VOX, N_DECONV[0], N_GRU_VOX, N_GRU_VOX)
        s = torch.zeros_like(h)
        gates = []
        residual = self.net_name != 'GRUNet'
        for t in range(T):                                    # theano.scan, sequential
            f = self.encoder_res(x[t]) if residual else self.encoder_gru(x[t])
            if self.net_name == 'ResidualLSTMNet':
                h, s, g = self.lstm_step(h, s, f)
            else:
                h, g = self.gru_step(h, f)
            gates.append(g)
        logits = self.decoder_res(h) if residual else self.decoder_gru(h)
        if y is None:
            yt = torch.zeros_like(logits)
        else:
            yt = torch.as_tensor(y).to(self.dtype)
        pred, loss = softmax_with_loss3d(logits, yt)
        return dict(pred=pred, loss=loss, update_gates=torch.stack(gates), logits=logits, h=h)

    def loss_and_grads(self, x, y):
        """models/net.py:56-58: tensor.grad(loss, params)."""
        for t in self.P:
            t.requires_grad_(True)
            t.grad = None
        out = self.forward(x, y)
        grads = torch.autograd.grad(out['loss'], self.P)
        return out, list(grads)


class RefAdam:
    """lib/solver.py:20-48 + 78-79,108: float32 iteration counter bumped BEFORE each call."""

    def __init__(self, net, lr=1e-4, w_decay=5e-5, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.net, self.lr, self.wd = net, lr, w_decay
        self.b1, self.b2, self.eps = beta_1, beta_2, epsilon
        self.t = 0
        self.m = [torch.zeros_like(p) for p in net.P]
        self.v = [torch.zeros_like(p) for p in net.P]

    def step(self, x, y):
        self.t += 1
        out, grads = self.net.loss_and_grads(x, y)
        t = float(self.t)
        lr_t = self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
        with torch.no_grad():
            for i, ((_, _, is_bias), p, g) in enumerate(zip(self.net.spec, self.net.P, grads)):
                rg = g if (is_bias or self.wd == 0) else g + self.wd * p
                self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * rg
                self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * rg * rg
                p -= lr_t * self.m[i] / (torch.sqrt(self.v[i]) + self.eps)
        return float(out['loss']), grads


def evaluate_voxel_prediction(preds, gt, thresh):
    """lib/voxel.py:4-11."""
    from .ref_numpy import evaluate_voxel_prediction as f
    return f(np.asarray(preds), np.asarray(gt), thresh)


def iou(pred, y, thresh=0.4):
    """IoU = intersection / union from lib/voxel.py:4-11 counts, summed over the batch
    (lib/test_net.py:62-66 loops samples)."""
    inter = union = 0
    for b in range(pred.shape[0]):
        r = evaluate_voxel_prediction(np.asarray(pred[b]), np.asarray(y[b]), thresh)
        inter += r[1]
        union += r[2]
    return inter / max(union, 1)
