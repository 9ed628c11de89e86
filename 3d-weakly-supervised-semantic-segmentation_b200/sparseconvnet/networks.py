"""scn.UNet / scn.FullyConvolutionalNet builders, composed from this package's modules.

Structure follows the reference's own mirrors of the upstream builders: UNet encoder half at
Function_test.py:113-164 (decoder half: BatchNorm + Deconvolution + JoinTable + blocks) and
FullyConvolutionalNet at Function_test.py:166-226; called at models/SparseConvNet.py:63-68,79-85,96-102.
"""
from . import modules as scn


def _block(dimension, residual_blocks, bn):
    def block(m, a, b):
        if residual_blocks:  # ResNet style blocks
            m.add(scn.ConcatTable()
                  .add(scn.Identity() if a == b else scn.NetworkInNetwork(a, b, False))
                  .add(scn.Sequential()
                       .add(bn(a))
                       .add(scn.SubmanifoldConvolution(dimension, a, b, 3, False))
                       .add(bn(b))
                       .add(scn.SubmanifoldConvolution(dimension, b, b, 3, False)))
                  ).add(scn.AddTable())
        else:  # VGG style blocks
            m.add(scn.Sequential()
                  .add(bn(a))
                  .add(scn.SubmanifoldConvolution(dimension, a, b, 3, False)))
    return block


def UNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2], leakiness=0, n_input_planes=-1):
    bn = lambda c: scn.BatchNormLeakyReLU(c, leakiness=leakiness)
    block = _block(dimension, residual_blocks, bn)

    def U(nPlanes, n_input_planes=-1):
        m = scn.Sequential()
        for i in range(reps):
            block(m, n_input_planes if n_input_planes != -1 else nPlanes[0], nPlanes[0])
            n_input_planes = -1
        if len(nPlanes) > 1:
            m.add(scn.ConcatTable().add(scn.Identity()).add(
                scn.Sequential()
                .add(bn(nPlanes[0]))
                .add(scn.Convolution(dimension, nPlanes[0], nPlanes[1], downsample[0], downsample[1], False))
                .add(U(nPlanes[1:]))
                .add(bn(nPlanes[1]))
                .add(scn.Deconvolution(dimension, nPlanes[1], nPlanes[0], downsample[0], downsample[1], False))))
            m.add(scn.JoinTable())
            for i in range(reps):
                block(m, nPlanes[0] * (2 if i == 0 else 1), nPlanes[0])
        return m
    return U(nPlanes, n_input_planes)


def FullyConvolutionalNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2]):
    block = _block(dimension, residual_blocks, lambda c: scn.BatchNormReLU(c))

    def U(nPlanes):
        m = scn.Sequential()
        for _ in range(reps):
            block(m, nPlanes[0], nPlanes[0])
        if len(nPlanes) > 1:
            m.add(scn.ConcatTable().add(scn.Identity()).add(
                scn.Sequential()
                .add(scn.BatchNormReLU(nPlanes[0]))
                .add(scn.Convolution(dimension, nPlanes[0], nPlanes[1], downsample[0], downsample[1], False))
                .add(U(nPlanes[1:]))
                .add(scn.UnPooling(dimension, downsample[0], downsample[1]))))
            m.add(scn.JoinTable())
        return m
    return U(nPlanes)
