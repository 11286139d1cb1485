"""torch-CPU restatement of the Pix2Pose encoder/decoder (both backbones).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, line by line:

* ``pix2pose_model/ae_model.py:175-240``  -- ``aemodel_unet_resnet50``
* ``pix2pose_model/ae_model.py:70-150``   -- ``aemodel_unet_prob`` ("paper" backbone)
* ``pix2pose_model/resnet50_mod.py:40-118, 200-213`` -- bottleneck blocks, stem

Semantics that are *not* torch defaults and are restated explicitly here:
TF ``padding='same'`` (asymmetric for k=5,s=2: 1 before / 2 after), Keras
``Conv2DTranspose(...,'same')`` = full scatter cropped ``[1:2N+1]``, inference
BatchNorm with epsilon 1e-3, ``LeakyReLU()`` alpha 0.3, NHWC ``Flatten`` order,
Keras kernel layouts (Conv2D ``(kh,kw,Cin,Cout)``, Conv2DTranspose ``(kh,kw,Cout,Cin)``,
Dense ``(in,out)``).

Weights come as a ``dict[str, np.ndarray]`` keyed by the canonical layer names of
``pix2pose_b200/weights.py`` (Keras layouts, fp32).
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # keras BatchNormalization default epsilon (ae_model.py:75, resnet50_mod.py:60)
LEAKY_ALPHA = 0.3    # keras LeakyReLU() default alpha (ae_model.py:78)


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _same_pad(n, k, s):
    """TF 'same' padding for one spatial dim -> (before, after, out)."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2, out


class NetOracle:
    """Functional forward of one backbone in NCHW internally, NHWC at the interface."""

    def __init__(self, weights, backbone="resnet50", dtype=torch.float32):
        assert backbone in ("resnet50", "paper")
        self.backbone = backbone
        self.dtype = dtype
        self.w = {k: _t(v, dtype) for k, v in weights.items()}
        self.taps = None  # optional dict filled with intermediate activations (NHWC numpy)

    # ---- primitive layers -------------------------------------------------------
    def conv(self, x, name, stride=1, padding="same"):
        k = self.w[name + "/kernel"]            # (kh,kw,Cin,Cout)
        b = self.w[name + "/bias"]
        kh, kw = k.shape[0], k.shape[1]
        if padding == "same":
            pt, pb, _ = _same_pad(x.shape[2], kh, stride)
            pl, pr, _ = _same_pad(x.shape[3], kw, stride)
            x = F.pad(x, (pl, pr, pt, pb))
        return F.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), b, stride=stride)

    def convT(self, x, name):
        """Conv2DTranspose(k=5, s=2, 'same'): full scatter (2N+3) cropped [1:2N+1]."""
        k = self.w[name + "/kernel"]            # (kh,kw,Cout,Cin)
        b = self.w[name + "/bias"]
        n_h, n_w = x.shape[2], x.shape[3]
        y = F.conv_transpose2d(x, k.permute(3, 2, 0, 1).contiguous(), b, stride=2, padding=1)
        return y[:, :, : 2 * n_h, : 2 * n_w]

    def bn(self, x, name):
        g, be = self.w[name + "/gamma"], self.w[name + "/beta"]
        mu, var = self.w[name + "/moving_mean"], self.w[name + "/moving_variance"]
        sc = g / torch.sqrt(var + BN_EPS)
        return x * sc.view(1, -1, 1, 1) + (be - mu * sc).view(1, -1, 1, 1)

    @staticmethod
    def lrelu(x):
        return F.leaky_relu(x, LEAKY_ALPHA)

    def dense(self, x, name):
        return x @ self.w[name + "/kernel"] + self.w[name + "/bias"]

    def _tap(self, name, x):
        if self.taps is not None:
            self.taps[name] = x.permute(0, 2, 3, 1).contiguous().numpy()

    # ---- resnet50_mod.py:40-118 -------------------------------------------------
    def _bottleneck(self, x, stage, block, stride, shortcut_conv):
        base = "res%d%s_branch" % (stage, block)
        bnb = "bn%d%s_branch" % (stage, block)
        y = self.conv(x, base + "2a", stride=stride, padding="valid")
        y = F.relu(self.bn(y, bnb + "2a"))
        y = self.conv(y, base + "2b", padding="same")
        y = F.relu(self.bn(y, bnb + "2b"))
        y = self.bn(self.conv(y, base + "2c", padding="valid"), bnb + "2c")
        if shortcut_conv:
            sc = self.bn(self.conv(x, base + "1", stride=stride, padding="valid"), bnb + "1")
        else:
            sc = x
        out = F.relu(y + sc)
        self._tap("act%d%s" % (stage, block), out)
        return out

    def _resnet_part(self, x):
        # resnet50_mod.py:200-204
        x = F.pad(x, (3, 3, 3, 3))
        x = self.conv(x, "conv1", stride=2, padding="valid")
        f1 = F.relu(self.bn(x, "bn_conv1"))
        self._tap("f1", f1)
        pt, pb, _ = _same_pad(f1.shape[2], 3, 2)
        pl, pr, _ = _same_pad(f1.shape[3], 3, 2)
        x = F.max_pool2d(F.pad(f1, (pl, pr, pt, pb), value=float("-inf")), 3, 2)
        self._tap("pool1", x)
        # :206-208
        x = self._bottleneck(x, 2, "a", 1, True)
        x = self._bottleneck(x, 2, "b", 1, False)
        f2 = self._bottleneck(x, 2, "c", 1, False)
        # :210-213
        x = self._bottleneck(f2, 3, "a", 2, True)
        x = self._bottleneck(x, 3, "b", 1, False)
        x = self._bottleneck(x, 3, "c", 1, False)
        f3 = self._bottleneck(x, 3, "d", 1, False)
        return f1, f2, f3

    def _twin(self, x, name):
        """conv{L}_1 / conv{L}_2 : Conv2D(5x5,s2,same)+BN+LeakyReLU twins (ae_model.py:74-106)."""
        a = self.lrelu(self.bn(self.conv(x, name + "_1", stride=2), "bn_" + name + "_1"))
        b = self.lrelu(self.bn(self.conv(x, name + "_2", stride=2), "bn_" + name + "_2"))
        return a, b

    # ---- forward ----------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x_nhwc):
        """x_nhwc: (N,128,128,3) array-like -> (decode (N,128,128,3), prob (N,128,128,1)) fp32 numpy."""
        x = _t(np.asarray(x_nhwc), self.dtype).permute(0, 3, 1, 2).contiguous()
        if self.backbone == "resnet50":
            f1, f2, f3 = self._resnet_part(x)
            s1, s2, s3 = f1[:, :32], f2[:, :128], f3[:, :128]     # ae_model.py:186-188
            enc_in = f3
        else:
            f1_1, f1_2 = self._twin(x, "conv1")
            f1 = torch.cat([f1_1, f1_2], 1)
            self._tap("f1", f1)
            f2_1, f2_2 = self._twin(f1, "conv2")
            f2 = torch.cat([f2_1, f2_2], 1)
            self._tap("f2", f2)
            f3_1, f3_2 = self._twin(f2, "conv3")
            enc_in = torch.cat([f3_1, f3_2], 1)
            self._tap("f3", enc_in)
            s1, s2, s3 = f1_2, f2_2, f3_2
        f4_1, f4_2 = self._twin(enc_in, "conv4")
        f4 = torch.cat([f4_1, f4_2], 1)                            # 8x8x512
        self._tap("f4", f4)
        flat = f4.permute(0, 2, 3, 1).reshape(f4.shape[0], -1)     # NHWC Flatten
        enc = self.dense(flat, "dense_1")
        d = self.dense(enc, "dense_2")
        d = d.reshape(-1, 8, 8, 256).permute(0, 3, 1, 2).contiguous()
        self._tap("d0", d)
        d = self.lrelu(self.bn(self.convT(d, "convT1"), "bn_convT1"))
        self._tap("d1", d)
        d = self.lrelu(self.bn(self.conv(torch.cat([d, s3], 1), "deconv1"), "bn_deconv1"))
        self._tap("d1_uni", d)
        d = self.lrelu(self.bn(self.convT(d, "convT2"), "bn_convT2"))
        self._tap("d2", d)
        d = self.lrelu(self.bn(self.conv(torch.cat([d, s2], 1), "deconv2"), "bn_deconv2"))
        self._tap("d2_uni", d)
        d = self.lrelu(self.bn(self.convT(d, "convT3"), "bn_convT3"))
        self._tap("d3", d)
        d = self.lrelu(self.bn(self.conv(torch.cat([d, s1], 1), "deconv3"), "bn_deconv3"))
        self._tap("d3_uni", d)
        dec = torch.tanh(self.convT(d, "convT_xyz"))
        prob = torch.sigmoid(self.convT(d, "convT_prob"))
        dec = dec.permute(0, 2, 3, 1).contiguous().to(torch.float32).numpy()
        prob = prob.permute(0, 2, 3, 1).contiguous().to(torch.float32).numpy()
        return dec, prob

    # keras duck type used by recognition.py:84,129
    def predict(self, x):
        d, p = self.forward(np.asarray(x, dtype=np.float32))
        return [d, p]
