"""The reference's `network_definition` bodies (models/gru_net.py:16-144,
models/res_gru_net.py:16-204) executed through the drop-in layer classes of lib/layers.py,
one C-ABI kernel call per layer, with the per-view encoder INSIDE the recurrence exactly as the
reference's theano.scan step has it.  The only textual change w.r.t. the reference is that
`theano` is lib/theano_compat (scan = host loop).

`LayerwiseGRUNet` / `LayerwiseResidualGRUNet` are ordinary user-defined networks in the sense of
models/net.py:43-51: `Net` runs them define-by-run, `Solver.train_loss` trains them (tensor.grad = the
tape the layers record).  They demonstrate that each layer is a drop-in and serve as a parity aid for
the fused engine; the shipped networks themselves run through engine.py.
"""
import numpy as np
import torch

from ..lib import layers as ly
from ..lib import theano_compat as theano
from ..lib.theano_compat import tensor
from ..lib.config import cfg
from ..lib.layers import TensorProductLayer, ConvLayer, PoolLayer, Unpool3DLayer, \
    LeakyReLU, SoftmaxWithLoss3D, Conv3DLayer, InputLayer, FlattenLayer, \
    FCConv3DLayer, TanhLayer, SigmoidLayer, ComplementLayer, AddLayer, \
    EltwiseMultiplyLayer, get_trainable_params
from .net import Net, tensor5


class LayerwiseResidualGRUNet(Net):
    RESIDUAL = True

    def network_definition(self):
        # (multi_views, self.batch_size, 3, self.img_h, self.img_w)
        self.x = tensor5()
        self.is_x_tensor4 = False
        _define(self, self.RESIDUAL)


class LayerwiseGRUNet(LayerwiseResidualGRUNet):
    RESIDUAL = False


_cache = {}


def run(net, x, y, residual):
    """Forward pass of an engine network's current parameters through the layer-by-layer definition."""
    compute = net.engine.compute
    key = (residual, net.batch_size, net.img_w, net.img_h, compute)
    lw = _cache.get(key)
    saved = (cfg.CONST.BATCH_SIZE, cfg.CONST.COMPUTE, cfg.CONST.IMG_W, cfg.CONST.IMG_H)
    cfg.CONST.BATCH_SIZE, cfg.CONST.COMPUTE = net.batch_size, compute
    cfg.CONST.IMG_W, cfg.CONST.IMG_H = net.img_w, net.img_h
    try:
        if lw is None:
            n0 = len(ly.trainable_params)
            lw = (LayerwiseResidualGRUNet if residual else LayerwiseGRUNet)()
            del ly.trainable_params[n0:]         # do not leak into the module-global list (quirk 13)
            _cache[key] = lw
        for w, v in zip(lw._weights, net.engine.get_params()):
            w.val.set_value(v)
        return lw.forward(x, y)
    finally:
        cfg.CONST.BATCH_SIZE, cfg.CONST.COMPUTE, cfg.CONST.IMG_W, cfg.CONST.IMG_H = saved


def _define(net, residual):
    batch_size, img_w, img_h = net.batch_size, net.img_w, net.img_h
    n_gru_vox = 4
    n_convfilter = [96, 128, 256, 256, 256, 256]
    n_fc_filters = [1024]
    n_deconvfilter = [128, 128, 128, 64, 32, 2]
    input_shape = (batch_size, 3, img_w, img_h)

    # ---- definition part: creates the Weights in the reference's order -----------------
    xl = InputLayer(input_shape)
    if residual:
        conv1a = ConvLayer(xl, (n_convfilter[0], 7, 7))
        conv1b = ConvLayer(conv1a, (n_convfilter[0], 3, 3))
        pool1 = PoolLayer(conv1b)
        conv2a = ConvLayer(pool1, (n_convfilter[1], 3, 3))
        conv2b = ConvLayer(conv2a, (n_convfilter[1], 3, 3))
        conv2c = ConvLayer(pool1, (n_convfilter[1], 1, 1))
        pool2 = PoolLayer(conv2c)
        conv3a = ConvLayer(pool2, (n_convfilter[2], 3, 3))
        conv3b = ConvLayer(conv3a, (n_convfilter[2], 3, 3))
        conv3c = ConvLayer(pool2, (n_convfilter[2], 1, 1))
        pool3 = PoolLayer(conv3b)
        conv4a = ConvLayer(pool3, (n_convfilter[3], 3, 3))
        conv4b = ConvLayer(conv4a, (n_convfilter[3], 3, 3))
        pool4 = PoolLayer(conv4b)
        conv5a = ConvLayer(pool4, (n_convfilter[4], 3, 3))
        conv5b = ConvLayer(conv5a, (n_convfilter[4], 3, 3))
        conv5c = ConvLayer(pool4, (n_convfilter[4], 1, 1))
        pool5 = PoolLayer(conv5b)
        conv6a = ConvLayer(pool5, (n_convfilter[5], 3, 3))
        conv6b = ConvLayer(conv6a, (n_convfilter[5], 3, 3))
        pool6 = PoolLayer(conv6b)
    else:
        conv1 = ConvLayer(xl, (n_convfilter[0], 7, 7))
        pool1 = PoolLayer(conv1)
        conv2 = ConvLayer(pool1, (n_convfilter[1], 3, 3))
        pool2 = PoolLayer(conv2)
        conv3 = ConvLayer(pool2, (n_convfilter[2], 3, 3))
        pool3 = PoolLayer(conv3)
        conv4 = ConvLayer(pool3, (n_convfilter[3], 3, 3))
        pool4 = PoolLayer(conv4)
        conv5 = ConvLayer(pool4, (n_convfilter[4], 3, 3))
        pool5 = PoolLayer(conv5)
        conv6 = ConvLayer(pool5, (n_convfilter[5], 3, 3))
        pool6 = PoolLayer(conv6)
    flat6 = FlattenLayer(pool6)
    fc7 = TensorProductLayer(flat6, n_fc_filters[0])
    s_shape = (batch_size, n_gru_vox, n_deconvfilter[0], n_gru_vox, n_gru_vox)
    prev_s = InputLayer(s_shape)
    t_x_s_update = FCConv3DLayer(prev_s, fc7, (n_deconvfilter[0], n_deconvfilter[0], 3, 3, 3))
    t_x_s_reset = FCConv3DLayer(prev_s, fc7, (n_deconvfilter[0], n_deconvfilter[0], 3, 3, 3))
    reset_gate = SigmoidLayer(t_x_s_reset)
    rs = EltwiseMultiplyLayer(reset_gate, prev_s)
    t_x_rs = FCConv3DLayer(rs, fc7, (n_deconvfilter[0], n_deconvfilter[0], 3, 3, 3))

    def recurrence(x_curr, prev_s_tensor, prev_in_gate_tensor):
        input_ = InputLayer(input_shape, x_curr)
        if residual:
            conv1a_ = ConvLayer(input_, (n_convfilter[0], 7, 7), params=conv1a.params)
            rect1a_ = LeakyReLU(conv1a_)
            conv1b_ = ConvLayer(rect1a_, (n_convfilter[0], 3, 3), params=conv1b.params)
            rect1_ = LeakyReLU(conv1b_)
            pool1_ = PoolLayer(rect1_)
            conv2a_ = ConvLayer(pool1_, (n_convfilter[1], 3, 3), params=conv2a.params)
            rect2a_ = LeakyReLU(conv2a_)
            conv2b_ = ConvLayer(rect2a_, (n_convfilter[1], 3, 3), params=conv2b.params)
            rect2_ = LeakyReLU(conv2b_)
            conv2c_ = ConvLayer(pool1_, (n_convfilter[1], 1, 1), params=conv2c.params)
            res2_ = AddLayer(conv2c_, rect2_)
            pool2_ = PoolLayer(res2_)
            conv3a_ = ConvLayer(pool2_, (n_convfilter[2], 3, 3), params=conv3a.params)
            rect3a_ = LeakyReLU(conv3a_)
            conv3b_ = ConvLayer(rect3a_, (n_convfilter[2], 3, 3), params=conv3b.params)
            rect3_ = LeakyReLU(conv3b_)
            conv3c_ = ConvLayer(pool2_, (n_convfilter[2], 1, 1), params=conv3c.params)
            res3_ = AddLayer(conv3c_, rect3_)
            pool3_ = PoolLayer(res3_)
            conv4a_ = ConvLayer(pool3_, (n_convfilter[3], 3, 3), params=conv4a.params)
            rect4a_ = LeakyReLU(conv4a_)
            conv4b_ = ConvLayer(rect4a_, (n_convfilter[3], 3, 3), params=conv4b.params)
            rect4_ = LeakyReLU(conv4b_)
            pool4_ = PoolLayer(rect4_)
            conv5a_ = ConvLayer(pool4_, (n_convfilter[4], 3, 3), params=conv5a.params)
            rect5a_ = LeakyReLU(conv5a_)
            conv5b_ = ConvLayer(rect5a_, (n_convfilter[4], 3, 3), params=conv5b.params)
            rect5_ = LeakyReLU(conv5b_)
            conv5c_ = ConvLayer(pool4_, (n_convfilter[4], 1, 1), params=conv5c.params)
            res5_ = AddLayer(conv5c_, rect5_)
            pool5_ = PoolLayer(res5_)
            conv6a_ = ConvLayer(pool5_, (n_convfilter[5], 3, 3), params=conv6a.params)
            rect6a_ = LeakyReLU(conv6a_)
            conv6b_ = ConvLayer(rect6a_, (n_convfilter[5], 3, 3), params=conv6b.params)
            rect6_ = LeakyReLU(conv6b_)
            res6_ = AddLayer(pool5_, rect6_)
            pool6_ = PoolLayer(res6_)
            flat6_ = FlattenLayer(pool6_)
        else:
            conv1_ = ConvLayer(input_, (n_convfilter[0], 7, 7), params=conv1.params)
            pool1_ = PoolLayer(conv1_)
            rect1_ = LeakyReLU(pool1_)
            conv2_ = ConvLayer(rect1_, (n_convfilter[1], 3, 3), params=conv2.params)
            pool2_ = PoolLayer(conv2_)
            rect2_ = LeakyReLU(pool2_)
            conv3_ = ConvLayer(rect2_, (n_convfilter[2], 3, 3), params=conv3.params)
            pool3_ = PoolLayer(conv3_)
            rect3_ = LeakyReLU(pool3_)
            conv4_ = ConvLayer(rect3_, (n_convfilter[3], 3, 3), params=conv4.params)
            pool4_ = PoolLayer(conv4_)
            rect4_ = LeakyReLU(pool4_)
            conv5_ = ConvLayer(rect4_, (n_convfilter[4], 3, 3), params=conv5.params)
            pool5_ = PoolLayer(conv5_)
            rect5_ = LeakyReLU(pool5_)
            conv6_ = ConvLayer(rect5_, (n_convfilter[5], 3, 3), params=conv6.params)
            pool6_ = PoolLayer(conv6_)
            rect6_ = LeakyReLU(pool6_)
            flat6_ = FlattenLayer(rect6_)
        fc7_ = TensorProductLayer(flat6_, n_fc_filters[0], params=fc7.params)
        rect7_ = LeakyReLU(fc7_)
        prev_s_ = InputLayer(s_shape, prev_s_tensor)
        t_x_s_update_ = FCConv3DLayer(prev_s_, rect7_, (n_deconvfilter[0], n_deconvfilter[0], 3, 3, 3),
                                      params=t_x_s_update.params)
        t_x_s_reset_ = FCConv3DLayer(prev_s_, rect7_, (n_deconvfilter[0], n_deconvfilter[0], 3, 3, 3),
                                     params=t_x_s_reset.params)
        update_gate_ = SigmoidLayer(t_x_s_update_)
        comp_update_gate_ = ComplementLayer(update_gate_)
        reset_gate_ = SigmoidLayer(t_x_s_reset_)
        rs_ = EltwiseMultiplyLayer(reset_gate_, prev_s_)
        t_x_rs_ = FCConv3DLayer(rs_, rect7_, (n_deconvfilter[0], n_deconvfilter[0], 3, 3, 3),
                                params=t_x_rs.params)
        tanh_t_x_rs_ = TanhLayer(t_x_rs_)
        gru_out_ = AddLayer(EltwiseMultiplyLayer(update_gate_, prev_s_),
                            EltwiseMultiplyLayer(comp_update_gate_, tanh_t_x_rs_))
        return gru_out_.output, update_gate_.output

    # ---- decoder definition (weights first so creation order matches the reference) -------
    gru_s = InputLayer(s_shape)          # placeholder, re-instantiated below with s_last
    if residual:
        unpool7 = Unpool3DLayer(gru_s)
        conv7a = Conv3DLayer(unpool7, (n_deconvfilter[1], 3, 3, 3))
        conv7b = Conv3DLayer(conv7a, (n_deconvfilter[1], 3, 3, 3))
        unpool8 = Unpool3DLayer(conv7b)
        conv8a = Conv3DLayer(unpool8, (n_deconvfilter[2], 3, 3, 3))
        conv8b = Conv3DLayer(conv8a, (n_deconvfilter[2], 3, 3, 3))
        unpool9 = Unpool3DLayer(conv8b)
        conv9a = Conv3DLayer(unpool9, (n_deconvfilter[3], 3, 3, 3))
        conv9b = Conv3DLayer(conv9a, (n_deconvfilter[3], 3, 3, 3))
        conv9c = Conv3DLayer(unpool9, (n_deconvfilter[3], 1, 1, 1))
        conv10a = Conv3DLayer(conv9c, (n_deconvfilter[4], 3, 3, 3))
        conv10b = Conv3DLayer(conv10a, (n_deconvfilter[4], 3, 3, 3))
        conv10c = Conv3DLayer(conv10a, (n_deconvfilter[4], 3, 3, 3))
        conv11 = Conv3DLayer(conv10c, (n_deconvfilter[5], 3, 3, 3))
    else:
        unpool7 = Unpool3DLayer(gru_s)
        conv7 = Conv3DLayer(unpool7, (n_deconvfilter[1], 3, 3, 3))
        unpool8 = Unpool3DLayer(conv7)
        conv8 = Conv3DLayer(unpool8, (n_deconvfilter[2], 3, 3, 3))
        unpool9 = Unpool3DLayer(conv8)
        conv9 = Conv3DLayer(unpool9, (n_deconvfilter[3], 3, 3, 3))
        conv10 = Conv3DLayer(conv9, (n_deconvfilter[4], 3, 3, 3))
        conv11 = Conv3DLayer(conv10, (n_deconvfilter[5], 3, 3, 3))

    s_update, _ = theano.scan(recurrence,
                              sequences=[net.x],
                              outputs_info=[tensor.zeros_like(np.zeros(s_shape), dtype=theano.config.floatX),
                                            tensor.zeros_like(np.zeros(s_shape), dtype=theano.config.floatX)])
    update_all = s_update[-1]
    s_all = s_update[0]
    s_last = s_all[-1]

    gru_s = InputLayer(s_shape, s_last)
    if residual:
        unpool7 = Unpool3DLayer(gru_s)
        conv7a_ = Conv3DLayer(unpool7, (n_deconvfilter[1], 3, 3, 3), params=conv7a.params)
        rect7a = LeakyReLU(conv7a_)
        conv7b_ = Conv3DLayer(rect7a, (n_deconvfilter[1], 3, 3, 3), params=conv7b.params)
        rect7 = LeakyReLU(conv7b_)
        res7 = AddLayer(unpool7, rect7)
        unpool8 = Unpool3DLayer(res7)
        conv8a_ = Conv3DLayer(unpool8, (n_deconvfilter[2], 3, 3, 3), params=conv8a.params)
        rect8a = LeakyReLU(conv8a_)
        conv8b_ = Conv3DLayer(rect8a, (n_deconvfilter[2], 3, 3, 3), params=conv8b.params)
        rect8 = LeakyReLU(conv8b_)
        res8 = AddLayer(unpool8, rect8)
        unpool9 = Unpool3DLayer(res8)
        conv9a_ = Conv3DLayer(unpool9, (n_deconvfilter[3], 3, 3, 3), params=conv9a.params)
        rect9a = LeakyReLU(conv9a_)
        conv9b_ = Conv3DLayer(rect9a, (n_deconvfilter[3], 3, 3, 3), params=conv9b.params)
        rect9 = LeakyReLU(conv9b_)
        conv9c_ = Conv3DLayer(unpool9, (n_deconvfilter[3], 1, 1, 1), params=conv9c.params)
        res9 = AddLayer(conv9c_, rect9)
        conv10a_ = Conv3DLayer(res9, (n_deconvfilter[4], 3, 3, 3), params=conv10a.params)
        rect10a = LeakyReLU(conv10a_)
        conv10b_ = Conv3DLayer(rect10a, (n_deconvfilter[4], 3, 3, 3), params=conv10b.params)
        rect10 = LeakyReLU(conv10b_)
        conv10c_ = Conv3DLayer(rect10a, (n_deconvfilter[4], 3, 3, 3), params=conv10c.params)
        res10 = AddLayer(conv10c_, rect10)
        conv11_ = Conv3DLayer(res10, (n_deconvfilter[5], 3, 3, 3), params=conv11.params)
    else:
        unpool7 = Unpool3DLayer(gru_s)
        conv7_ = Conv3DLayer(unpool7, (n_deconvfilter[1], 3, 3, 3), params=conv7.params)
        rect7 = LeakyReLU(conv7_)
        unpool8 = Unpool3DLayer(rect7)
        conv8_ = Conv3DLayer(unpool8, (n_deconvfilter[2], 3, 3, 3), params=conv8.params)
        rect8 = LeakyReLU(conv8_)
        unpool9 = Unpool3DLayer(rect8)
        conv9_ = Conv3DLayer(unpool9, (n_deconvfilter[3], 3, 3, 3), params=conv9.params)
        rect9 = LeakyReLU(conv9_)
        conv10_ = Conv3DLayer(rect9, (n_deconvfilter[4], 3, 3, 3), params=conv10.params)
        rect10 = LeakyReLU(conv10_)
        conv11_ = Conv3DLayer(rect10, (n_deconvfilter[5], 3, 3, 3), params=conv11.params)
    softmax_loss = SoftmaxWithLoss3D(conv11_.output)
    net.loss = softmax_loss.loss(net.y)
    net.error = softmax_loss.error(net.y)
    net.params = get_trainable_params()
    net.output = softmax_loss.prediction()
    net.activations = [update_all]
