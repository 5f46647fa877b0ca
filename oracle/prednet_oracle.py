"""ORACLE (test infrastructure only) -- fp32 CPU restatement of the reference PredNet layer.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  The product (tezip_b200/) never does.

Follows /root/reference/src/prednet.py:
  * weight list order  ............ prednet.py:210-227  (sorted keys a, ahat, c, f, i, o; layer ascending;
                                     [kernel(kh,kw,Cin,Cout), bias(Cout)] per conv)
  * zero initial r, c, e states .... prednet.py:143-190
  * one time step .................. prednet.py:235-308
  * Model.predict protocol ......... compress.py:191-197,218-229 ; decompress.py:141-143,150-178
    (fresh zero state on every call, return_sequences=True, output_mode='prediction')

Third-party arithmetic that is NOT in /root/reference (parity unpinned): the reference runs
these ops through tensorflow-gpu 1.15 / keras 2.2.4 / cuDNN 7.6.5 (docs/index.rst:263-264,188).  Restated
here from the published Keras 2.2.4 semantics: Conv2D 'same' stride 1 cross-correlation with HWIO kernels,
MaxPooling2D 2x2/2 'valid', UpSampling2D nearest 2x, hard_sigmoid(x) = clip(0.2*x + 0.5, 0, 1), tanh, relu.
"""
import numpy as np
import torch
import torch.nn.functional as F

CONV_KEYS_SORTED = ("a", "ahat", "c", "f", "i", "o")  # prednet.py:212 sorted(self.conv_layers.keys())


def conv_specs(stack_sizes, R_stack_sizes):
    """(key, layer, Cin, Cout) in Keras weight-list order (prednet.py:212-227)."""
    L = len(stack_sizes)
    out = []
    for c in CONV_KEYS_SORTED:
        n_l = L - 1 if c == "a" else L  # prednet.py:204-205: 'a' convs exist for l < L-1 only
        for l in range(n_l):
            if c == "ahat":
                cin, cout = R_stack_sizes[l], stack_sizes[l]          # prednet.py:215-216, 202
            elif c == "a":
                cin, cout = 2 * stack_sizes[l], stack_sizes[l + 1]    # prednet.py:217-218, 205
            else:
                cin = 2 * stack_sizes[l] + R_stack_sizes[l]           # prednet.py:220
                if l < L - 1:
                    cin += R_stack_sizes[l + 1]                       # prednet.py:221-222
                cout = R_stack_sizes[l]                               # prednet.py:199
            out.append((c, l, cin, cout))
    return out


def hard_sigmoid(x):
    # keras 2.2.4 tensorflow_backend.hard_sigmoid: clip(0.2*x + 0.5, 0, 1)
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


class PredNetOracle:
    """output_mode='prediction', extrap_start_time=None, channels_last (the only mode compress/decompress use:
    compress.py:164, decompress.py:76)."""

    def __init__(self, weights, stack_sizes, R_stack_sizes, pixel_max=1.0):
        self.stack = tuple(int(s) for s in stack_sizes)
        self.R = tuple(int(s) for s in R_stack_sizes)
        self.L = len(self.stack)
        self.pixel_max = float(pixel_max)
        specs = conv_specs(self.stack, self.R)
        assert len(weights) == 2 * len(specs), (len(weights), len(specs))
        self.w = {}
        for n, (c, l, cin, cout) in enumerate(specs):
            k = np.asarray(weights[2 * n], dtype=np.float32)
            b = np.asarray(weights[2 * n + 1], dtype=np.float32)
            assert k.shape == (3, 3, cin, cout), (c, l, k.shape, (3, 3, cin, cout))
            assert b.shape == (cout,)
            # HWIO -> OIHW for torch; both frameworks compute cross-correlation (no kernel flip)
            self.w[(c, l)] = (torch.from_numpy(np.ascontiguousarray(k.transpose(3, 2, 0, 1))),
                              torch.from_numpy(b.copy()))

    def _conv(self, key, l, x):
        k, b = self.w[(key, l)]
        return F.conv2d(x, k, b, stride=1, padding=1)  # Conv2D(padding='same') 3x3

    def zero_state(self, B, Hp, Wp):
        # prednet.py:143-190 -- all-zero r, c (R[l] channels) and e (2*stack[l]) at Hp/2^l x Wp/2^l
        r = [torch.zeros(B, self.R[l], Hp >> l, Wp >> l) for l in range(self.L)]
        c = [torch.zeros(B, self.R[l], Hp >> l, Wp >> l) for l in range(self.L)]
        e = [torch.zeros(B, 2 * self.stack[l], Hp >> l, Wp >> l) for l in range(self.L)]
        return r, c, e

    def step(self, a, r_tm1, c_tm1, e_tm1):
        """prednet.py:235-308.  a: [B,C,Hp,Wp] (NCHW here; channel concat order is what matters)."""
        L = self.L
        c_new = [None] * L
        r_new = [None] * L
        e_new = []
        r_up = None
        for l in reversed(range(L)):                                   # prednet.py:249
            inputs = [r_tm1[l], e_tm1[l]]                              # prednet.py:250
            if l < L - 1:
                inputs.append(r_up)                                    # prednet.py:251-252
            x = torch.cat(inputs, dim=1)                               # prednet.py:254
            i = hard_sigmoid(self._conv("i", l, x))                    # prednet.py:255
            f = hard_sigmoid(self._conv("f", l, x))                    # prednet.py:256
            o = hard_sigmoid(self._conv("o", l, x))                    # prednet.py:257
            _c = f * c_tm1[l] + i * torch.tanh(self._conv("c", l, x))  # prednet.py:258
            _r = o * torch.tanh(_c)                                    # prednet.py:259
            c_new[l] = _c
            r_new[l] = _r
            if l > 0:
                r_up = F.interpolate(_r, scale_factor=2, mode="nearest")  # prednet.py:263-264
        frame_prediction = None
        for l in range(L):                                             # prednet.py:267
            ahat = torch.relu(self._conv("ahat", l, r_new[l]))         # prednet.py:268 (+ :201-202)
            if l == 0:
                ahat = torch.clamp(ahat, max=self.pixel_max)           # prednet.py:269-270
                frame_prediction = ahat                                # prednet.py:271
            e_up = torch.relu(ahat - a)                                # prednet.py:274
            e_down = torch.relu(a - ahat)                              # prednet.py:275
            e_new.append(torch.cat((e_up, e_down), dim=1))             # prednet.py:277
            if l < L - 1:
                a = torch.relu(self._conv("a", l, e_new[l]))           # prednet.py:290
                a = F.max_pool2d(a, 2, 2)                              # prednet.py:291
        return frame_prediction, r_new, c_new, e_new                   # prednet.py:293-295,305

    @torch.no_grad()
    def predict(self, x, batch_size=None):
        """keras Model.predict on x[B,T,Hp,Wp,C] -> [B,T,Hp,Wp,C] float32; zero state at every call."""
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))  # keras casts to floatx
        B, T, Hp, Wp, C = x.shape
        assert C == self.stack[0] and Hp % (1 << (self.L - 1)) == 0 and Wp % (1 << (self.L - 1)) == 0
        r, c, e = self.zero_state(B, Hp, Wp)
        outs = []
        for t in range(T):
            a = x[:, t].permute(0, 3, 1, 2).contiguous()
            p, r, c, e = self.step(a, r, c, e)
            outs.append(p.permute(0, 2, 3, 1))
        return torch.stack(outs, dim=1).contiguous().numpy()

    # Convenience forms of the call protocol
    def next(self, frames):
        """frames[B,Hp,Wp,C] float32 -> Model.predict([frame, zeros])[:,1]  (compress.py:224-229)."""
        frames = np.asarray(frames, dtype=np.float32)
        x = np.stack([frames, np.zeros_like(frames)], axis=1)
        return self.predict(x)[:, 1]

    def p0(self, Hp, Wp):
        """Model.predict(anything)[0,0]: input independent (compress.py:197; decompress.py:143)."""
        x = np.zeros((1, 1, Hp, Wp, self.stack[0]), np.float32)
        return self.predict(x)[0, 0]
