"""Architecture tables of the 2d3d-ResNet backbones on the hot path.

Restated from /root/reference/backbone/resnet_2d3d.py:205-284 (ResNet2d3d_full, _make_layer,
resnet18/34_2d3d_full) and backbone/select_backbone.py:3-21.  Bottleneck networks (resnet50+,
resnet_2d3d.py:286-308) are outside SURVEY.md §8's scope and raise NotImplementedError.
"""
NETWORKS = {
    'resnet18': (2, 2, 2, 2),     # blocks per stage; block types [2d, 2d, 3d, 3d]
    'resnet34': (3, 4, 6, 3),
}
UNSUPPORTED = ('resnet50', 'resnet101', 'resnet152', 'resnet200')
STAGE_PLANES = (64, 128, 256, 256)       # layer4 narrowed to 256 planes (resnet_2d3d.py:222)
STAGE_IS3D = (False, False, True, True)
FEATURE_SIZE = 256                        # select_backbone.py:7,10


def backbone_spec(network):
    if network in UNSUPPORTED:
        raise NotImplementedError('%s (Bottleneck blocks) is outside the B200 hot-path scope' % network)
    if network not in NETWORKS:
        raise IOError('model type is wrong')               # select_backbone.py:19
    spec = []
    inplanes = 64
    for si, nblocks in enumerate(NETWORKS[network]):
        planes = STAGE_PLANES[si]
        for bi in range(nblocks):
            s = (1 if si == 0 else 2) if bi == 0 else 1
            ds = (bi == 0) and (s != 1 or inplanes != planes)          # resnet_2d3d.py:234
            last = (si == 3 and bi == nblocks - 1)
            spec.append(dict(name='layer%d.%d' % (si + 1, bi), stage=si + 1, index=bi, inplanes=inplanes,
                             planes=planes, stride=s, is3d=STAGE_IS3D[si], downsample=ds,
                             final_relu=not last))                     # resnet_2d3d.py:249-252
            inplanes = planes
    return spec
