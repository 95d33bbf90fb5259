"""TEST INFRASTRUCTURE ONLY -- ``add_export_config`` of detectron2 v0.5 only adds export-time keys the path never reads."""


def add_export_config(cfg):
    return cfg
