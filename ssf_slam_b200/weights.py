"""Deterministic random-init weights in the reference's ``state_dict`` format.

The trained checkpoint is not in the reference tree (``.MISSING_LARGE_BLOBS``: ``pretrained_model/model.best.t7``) and
BASELINE.json's configs ask for random-init weights, so bench / smoke / tests share this generator.  Keys and shapes are
exactly those of ``TFlowV3_Occlussion.TFlow().state_dict()`` (317 tensors; checked by a strict load into the unmodified
reference in oracle/gen_golden.py).  BatchNorm running statistics are non-trivial on purpose so that folding is exercised.
"""
import torch


def random_init_state_dict(seed=0, flow_channels=3, input_channels=3):
    """State dict with the reference's key names and shapes, filled deterministically WITHOUT the
    reference (for the GPU box).  ``flow_channels=4`` gives the ``add_Seg_after_FLow=True`` variant's 4-channel flow
    heads (ASF/utils/soflow.py:343-346); ``input_channels=4`` the ``_afterPC`` variant's first layer
    (ASF/TFlowV3_Occlussion_addSeg_afterPC.py:68).  NOT the reference's init distribution draw-for-draw: goldens that
    must match the reference use the state_dict saved by oracle/gen_golden.py instead."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cout, cin, dims, bias):
        bound = 1.0 / (cin ** 0.5)
        sd[key + ".weight"] = (torch.rand((cout, cin) + (1,) * dims, generator=g) * 2 - 1) * bound
        if bias:
            sd[key + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(key, c):
        sd[key + ".weight"] = 1.0 + 0.1 * (torch.rand(c, generator=g) * 2 - 1)
        sd[key + ".bias"] = 0.1 * (torch.rand(c, generator=g) * 2 - 1)
        sd[key + ".running_mean"] = 0.1 * (torch.rand(c, generator=g) * 2 - 1)
        sd[key + ".running_var"] = 1.0 + 0.2 * torch.rand(c, generator=g)
        sd[key + ".num_batches_tracked"] = torch.tensor(0)

    conv("point_conv.0.composed_module.0", 32, input_channels, 1, False)
    conv("point_conv.1.composed_module.0", 32, 32, 1, False)
    for name, cin, mlp in (("sa1", 32, [32, 32, 64]), ("sa2", 64, [64, 64, 128]),
                           ("sa3", 128, [128, 128, 256]), ("sa4", 256, [256, 256, 512])):
        last = cin + 3
        for i, c in enumerate(mlp):
            conv("%s.mlp_convs.%d" % (name, i), c, last, 2, False)
            last = c
        for i, c in enumerate(mlp):
            bn("%s.mlp_bns.%d" % (name, i), c)
    for name, c1, c2, mlp, mlp2 in (("su3", 256, 512, [256, 256], [256, 256]), ("su2", 128, 256, [128, 128], [128, 128]),
                                    ("su1", 64, 128, [64, 64], [64, 64]), ("su0", 32, 64, [64, 64], [64, 64])):
        last = c2 + 3
        for i, c in enumerate(mlp):
            conv("%s.mlp1_convs.%d.0" % (name, i), c, last, 2, False)
            bn("%s.mlp1_convs.%d.1" % (name, i), c)
            last = c
        last = mlp[-1] + c1
        for i, c in enumerate(mlp2):
            conv("%s.mlp2_convs.%d.0" % (name, i), c, last, 1, False)
            bn("%s.mlp2_convs.%d.1" % (name, i), c)
            last = c
    for name, cin, sfc, mlp, fmlp in (("flow3_r", 256, 0, [256, 256], [128, 128]), ("flow2_r", 192, 128, [128, 128], [128, 128]),
                                      ("flow1_r", 96, 128, [64, 64], [64, 64]), ("flow0_r", 96, 64, [64, 64], [64, 64])):
        p = name + ".cost"
        for grp in ("mlp_convs", "mlp_convs2"):
            last = cin * 2
            for i, c in enumerate(mlp):
                conv("%s.%s.%d" % (p, grp, i), c, last, 2, True)
                last = c
        m = mlp[-1]
        conv(p + ".weightnet1.0", m, m, 2, False)
        bn(p + ".weightnet1.1", m)
        conv(p + ".weightnet1.3", m // 2, m, 2, False)
        bn(p + ".weightnet1.4", m // 2)
        conv(p + ".weightnet1.6", 1, m // 2, 2, True)
        last = m + sfc + 3
        for i, c in enumerate(mlp):
            conv("%s.mlp_convs3.%d" % (p, i), c, last, 2, True)
            last = c
        last = m * 2 + sfc + 3
        for i, c in enumerate(mlp):
            conv("%s.mlp_convs4.%d" % (p, i), c, last, 2, True)
            last = c
        last = m
        for i, c in enumerate(fmlp):
            conv("%s.flow_mlp_convs.%d.composed_module.0" % (p, i), c, last, 1, True)
            last = c
        conv(p + ".fc", flow_channels, last, 1, True)
    conv("deconv3_2.composed_module.0", 64, 256, 1, False)
    conv("deconv2_1.composed_module.0", 32, 128, 1, False)
    conv("deconv1_0.composed_module.0", 32, 64, 1, False)
    return sd
