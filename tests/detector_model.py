"""Synthetic SSD graphs in ncnn's text/bin format for the detector tests: same layer vocabulary and export idioms as the reference model
(Thirdparty/ncnn_model/mobilenetv3_ssdlite_voc.param: hard-swish as add/clip/mul/div with MemoryData scalars, spatial SE tail, residual adds,
SSDLite heads, mmdetection-style PriorBox, DetectionOutput), random weights, a few thousand priors."""
import os
import struct

import numpy as np


class Graph:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.layers = []      # [type, name, ins, outs, param string, [arrays]]
        self.consts = {}
        self.uid = 0

    def name(self, p='b'):
        self.uid += 1
        return '%s%d' % (p, self.uid)

    def add(self, typ, ins, params='', weights=(), nout=1, name=None):
        name = name or self.name('l')
        outs = [name] if nout == 1 else ['%s_%d' % (name, i) for i in range(nout)]
        self.layers.append([typ, name, list(ins), outs, params, list(weights)])
        return outs[0] if nout == 1 else outs

    def const(self, v):
        n = self.name('c')
        self.layers.insert(1, ['MemoryData', n, [], [n], '0=1', [np.array([v], np.float32)]])
        return n

    def conv(self, x, cin, cout, k=1, s=1, p=0, gain=1.0, dw=False):
        fan = k * k * (1 if dw else cin)
        w = (self.rng.standard_normal((cout, 1 if dw else cin, k, k)) * gain * np.sqrt(2.0 / fan)).astype(np.float32)
        b = (self.rng.standard_normal(cout) * 0.1).astype(np.float32)
        prm = '0=%d 1=%d 11=%d 2=1 12=1 3=%d 13=%d 4=%d 14=%d 5=1 6=%d' % (cout, k, k, s, s, p, p, w.size)
        if dw:
            prm += ' 7=%d' % cout
        return self.add('ConvolutionDepthWise' if dw else 'Convolution', [x], prm, [np.zeros(1, np.uint32), w, b])

    def relu(self, x): return self.add('ReLU', [x])
    def clip(self, x): return self.add('Clip', [x], '0=0.000000 1=6.000000')
    def binop(self, a, b, op): return self.add('BinaryOp', [a, b], '0=%d' % op)

    def hswish(self, x):
        return self.binop(self.binop(x, self.clip(self.binop(x, self.const(3.0), 0)), 2), self.const(6.0), 3)

    def hsigmoid(self, x):
        return self.binop(self.clip(self.binop(x, self.const(3.0), 0)), self.const(6.0), 3)

    def finalize(self):
        """ncnn graphs are single-consumer: insert Split layers."""
        out = []
        use = {}
        for L in self.layers:
            for i in L[2]:
                use[i] = use.get(i, 0) + 1
        renamed = {}
        for L in self.layers:
            ins = []
            for i in L[2]:
                if i in renamed:
                    ins.append(renamed[i].pop())
                else:
                    ins.append(i)
            L = [L[0], L[1], ins, L[3], L[4], L[5]]
            out.append(L)
            for o in L[3]:
                if use.get(o, 0) > 1:
                    names = ['%s_splitncnn_%d' % (o, q) for q in range(use[o])]
                    out.append(['Split', 'splitncnn_' + o, [o], names, '', []])
                    renamed[o] = names[:]
        self.layers = out

    def write(self, param_path, bin_path):
        self.finalize()
        nblobs = sum(len(L[3]) for L in self.layers)
        with open(param_path, 'w') as f:
            f.write('7767517\n%d %d\n' % (len(self.layers), nblobs))
            for typ, name, ins, outs, prm, _ in self.layers:
                f.write('%-24s %-24s %d %d %s %s\n' % (typ, name, len(ins), len(outs), ' '.join(ins + outs), prm))
        with open(bin_path, 'wb') as f:
            for L in self.layers:
                for a in L[5]:
                    f.write(np.ascontiguousarray(a).tobytes())


def write_mini_model(dirpath, seed=0, conf_gain=1.0, person_bias=2.0, many_priors=False):
    g = Graph(seed)
    data = g.add('Input', [], name='input')
    x = g.hswish(g.conv(data, 3, 8, 3, 2, 1))                                     # 8 x 150 x 150
    x = g.relu(g.conv(x, 8, 16)); x = g.relu(g.conv(x, 16, 16, 3, 2, 1, dw=True))  # 75
    x0 = g.conv(x, 16, 8)
    # SE + residual block
    y = g.hswish(g.conv(x0, 8, 24)); y = g.hswish(g.conv(y, 24, 24, 5, 1, 2, dw=True)); y = g.conv(y, 24, 8)
    se = g.conv(g.relu(g.conv(y, 8, 6)), 6, 8)
    x = g.binop(g.binop(y, g.hsigmoid(se), 2), x0, 0)
    # plain residual
    r = g.conv(g.relu(g.conv(g.relu(g.conv(x, 8, 16)), 16, 16, 3, 1, 1, dw=True)), 16, 8)
    x = g.binop(r, x, 0)
    x = g.clip(g.conv(g.clip(g.conv(x, 8, 8, 5, 2, 2, dw=True)), 8, 32))          # 38
    f1 = g.clip(g.conv(g.clip(g.conv(x, 32, 32, 3, 2, 1, dw=True)), 32, 32))     # 19
    f2 = g.clip(g.conv(g.clip(g.conv(g.clip(g.conv(f1, 32, 16)), 16, 16, 3, 2, 1, dw=True)), 16, 32))   # 10
    ncls = 21
    locs, confs, priors = [], [], []
    specs = [(x if many_priors else f1, 32, '-23300=1,60.000000 -23301=1,105.0 -23302=1,2.000000', 4), (f2, 32, '-23300=1,105.000000 -23301=1,150.0 -23302=2,2.000000,3.0', 6)]
    for k, (f, c, pb, npr) in enumerate(specs):
        cf = g.conv(g.clip(g.conv(f, c, c, 3, 1, 1, dw=True)), c, npr * ncls, gain=conf_gain)
        g.layers[-1][5][2][15::ncls] += np.float32(person_bias)            # make the person class (id 15) show up among the detections
        confs.append(g.add('Flatten', [g.add('Permute', [cf], '0=3')]))
        lc = g.conv(g.clip(g.conv(f, c, c, 3, 1, 1, dw=True)), c, npr * 4)
        locs.append(g.add('Flatten', [g.add('Permute', [lc], '0=3')]))
        priors.append(g.add('PriorBox', [f, data], pb + ' 3=0.100000 4=0.100000 5=0.200000 6=0.200000 7=1 8=0 9=-233 10=-233 11=-233.000000 12=-233.000000 13=0.500000 14=1 15=1'))
    loc = g.add('Concat', locs, '0=0', name='mbox_loc')
    conf = g.add('Concat', confs, '0=0', name='mbox_conf')
    pri = g.add('Concat', priors, '0=1', name='mbox_priorbox')
    sm = g.add('Softmax', [g.add('Reshape', [conf], '0=%d 1=-1' % ncls)], '0=1 1=1')
    g.add('DetectionOutput', [loc, g.add('Flatten', [sm]), pri], '0=%d 1=0.450000 2=300 3=100 4=0.010000' % ncls, name='detection_out')
    os.makedirs(dirpath, exist_ok=True)
    pp, bp = os.path.join(dirpath, 'mini.param'), os.path.join(dirpath, 'mini.bin')
    g.write(pp, bp)
    return pp, bp


def synthetic_rgb(h, w, seed):
    rng = np.random.default_rng(seed)
    img = np.full((h, w, 3), 110, np.float32) + rng.normal(0, 6, (h, w, 3))
    for _ in range(60):
        x0, y0 = rng.integers(0, w - 8), rng.integers(0, h - 8)
        x1, y1 = min(w, x0 + rng.integers(8, w // 3)), min(h, y0 + rng.integers(8, h // 3))
        img[y0:y1, x0:x1] = rng.integers(0, 256, 3)
    return np.clip(img, 0, 255).astype(np.uint8)
