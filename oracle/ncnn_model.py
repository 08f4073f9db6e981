"""TEST INFRASTRUCTURE: reader of the ncnn text graph (.param) and weight blob (.bin) of the detector
(Thirdparty/ncnn_model/mobilenetv3_ssdlite_voc.{param,bin}, loaded at src/Detector2D.cc:25-26).  ncnn is not vendored in the reference tree:
the format is restated from its public documentation (param: magic 7767517, "layers blobs", then one line per layer
"type name n_in n_out inputs... outputs... key=value..."; array keys are written as -233xx=count,v0,v1,...; bin: for every layer with weights,
in graph order, a 4-byte storage tag (0 = raw float32) in front of the weight tensor of Convolution / ConvolutionDepthWise, followed by the
bias floats; MemoryData stores raw float32 without a tag)."""
import numpy as np


class Layer:
    def __init__(self, typ, name, inputs, outputs, params):
        self.type, self.name, self.inputs, self.outputs, self.params = typ, name, inputs, outputs, params
        self.weight = None; self.bias = None; self.data = None

    def p(self, key, default=0):
        return self.params.get(key, default)


def parse_param(path):
    lines = [l.strip() for l in open(path) if l.strip()]
    assert lines[0] == '7767517', 'not an ncnn text param file'
    nlayers, nblobs = map(int, lines[1].split())
    layers = []
    for line in lines[2:]:
        tok = line.split()
        typ, name, nin, nout = tok[0], tok[1], int(tok[2]), int(tok[3])
        ins = tok[4:4 + nin]; outs = tok[4 + nin:4 + nin + nout]
        params = {}
        for kv in tok[4 + nin + nout:]:
            k, v = kv.split('=')
            k = int(k)
            if k <= -23300:                      # array: -23300 - id = count, values...
                vals = v.split(',')
                params[-(k + 23300)] = [float(x) for x in vals[1:1 + int(vals[0])]]
            else:
                params[k] = float(v) if ('.' in v or 'e' in v.lower()) else int(v)
        layers.append(Layer(typ, name, ins, outs, params))
    assert len(layers) == nlayers, (len(layers), nlayers)
    return layers


def load_weights(layers, bin_path):
    """Attaches weight / bias / constant arrays to the layers; returns the number of bytes consumed (must equal the file size)."""
    buf = np.fromfile(bin_path, np.uint8)
    off = 0

    def take_f32(n):
        nonlocal off
        a = buf[off:off + 4 * n].view(np.float32).copy(); off += 4 * n
        return a
    for L in layers:
        if L.type in ('Convolution', 'ConvolutionDepthWise'):
            tag = int(buf[off:off + 4].view(np.uint32)[0]); off += 4
            assert tag == 0, 'only raw float32 weights are expected (tag %#x)' % tag
            cout, k, wsize = L.p(0), L.p(1), L.p(6)
            group = L.p(7, 1) if L.type == 'ConvolutionDepthWise' else 1
            cin_g = wsize // (cout * k * k)
            L.weight = take_f32(wsize).reshape(cout, cin_g, k, k)
            L.group = group
            if L.p(5):
                L.bias = take_f32(cout)
        elif L.type == 'MemoryData':
            w, h, c = L.p(0), L.p(1, 0), L.p(2, 0)
            n = w * max(h, 1) * max(c, 1)
            L.data = take_f32(n)
    return off, len(buf)
