"""Backbone selector: the entry point the reference's models import
(/root/reference/backbone/select_backbone.py:3-21, used at dpc/model_3d.py:28, eval/model_3d_lc.py:28)."""
from .resnet_2d3d import (resnet18_2d3d_full, resnet34_2d3d_full, resnet50_2d3d_full, resnet101_2d3d_full,
                          resnet152_2d3d_full, resnet200_2d3d_full)

_FACTORIES = {
    'resnet18': (resnet18_2d3d_full, 256),
    'resnet34': (resnet34_2d3d_full, 256),
    'resnet50': (resnet50_2d3d_full, 1024),
    'resnet101': (resnet101_2d3d_full, 1024),
    'resnet152': (resnet152_2d3d_full, 1024),
    'resnet200': (resnet200_2d3d_full, 1024),
}


def select_resnet(network, track_running_stats=True):
    """-> (backbone module, {'feature_size': int}); unknown names raise IOError like the reference (:19)"""
    if network not in _FACTORIES:
        raise IOError('model type is wrong')
    factory, feature_size = _FACTORIES[network]
    model = factory(track_running_stats=track_running_stats)
    return model, {'feature_size': feature_size}
