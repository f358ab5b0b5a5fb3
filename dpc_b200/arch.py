"""Architecture tables of the 2d3d-ResNet backbones on the hot path.

Restated from /root/reference/backbone/resnet_2d3d.py:205-308 (ResNet2d3d_full, _make_layer, the
resnet18 ... resnet200 factories) and backbone/select_backbone.py:3-21.  r18 / r34 use BasicBlock2d/3d
(two 3x3 convs), r50+ Bottleneck2d/3d (1x1x1 -> 3x3 -> 1x1x1, expansion 4; resnet_2d3d.py:119-202).
"""
NETWORKS = {
    # name: (blocks per stage, block kind); block types are [2d, 2d, 3d, 3d] for all of them
    'resnet18': ((2, 2, 2, 2), 'basic'),
    'resnet34': ((3, 4, 6, 3), 'basic'),
    'resnet50': ((3, 4, 6, 3), 'bottleneck'),
    'resnet101': ((3, 4, 23, 3), 'bottleneck'),
    'resnet152': ((3, 8, 36, 3), 'bottleneck'),
    'resnet200': ((3, 24, 36, 3), 'bottleneck'),
}
STAGE_PLANES = (64, 128, 256, 256)       # layer4 narrowed to 256 planes (resnet_2d3d.py:222)
STAGE_IS3D = (False, False, True, True)
FEATURE_SIZE = 256                        # select_backbone.py:7,10 (BasicBlock networks)


def feature_size(network):
    """select_backbone.py:4-10: 256 for r18 / r34, 1024 (256 planes x expansion 4) for the Bottleneck networks"""
    if network not in NETWORKS:
        raise IOError('model type is wrong')               # select_backbone.py:19
    return 1024 if NETWORKS[network][1] == 'bottleneck' else FEATURE_SIZE


def backbone_spec(network):
    if network not in NETWORKS:
        raise IOError('model type is wrong')               # select_backbone.py:19
    layers, kind = NETWORKS[network]
    expansion = 4 if kind == 'bottleneck' else 1           # resnet_2d3d.py:48,84,120,162
    spec = []
    inplanes = 64
    for si, nblocks in enumerate(layers):
        planes = STAGE_PLANES[si]
        for bi in range(nblocks):
            s = (1 if si == 0 else 2) if bi == 0 else 1
            ds = (bi == 0) and (s != 1 or inplanes != planes * expansion)        # resnet_2d3d.py:234
            last = (si == 3 and bi == nblocks - 1)
            spec.append(dict(name='layer%d.%d' % (si + 1, bi), stage=si + 1, index=bi, block=kind, inplanes=inplanes,
                             planes=planes, outplanes=planes * expansion, stride=s, is3d=STAGE_IS3D[si],
                             downsample=ds, final_relu=not last))                # resnet_2d3d.py:249-252
            inplanes = planes * expansion
    return spec
