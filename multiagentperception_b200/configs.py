"""The reference's YAML config surface as plain dicts (keys and value conventions of configs/**/*.yml).

The hot path reads model.{arch, agent_num, enc_backbone, dec_backbone, feat_squeezer, feat_channel,
shuffle_features, attention, query, sparse, shared_img_encoder, query_size, key_size} and data.img_rows
(ptsemseg/models/__init__.py:13-84); everything else in the YAMLs is trainer-only. Shipped files can be loaded as-is
with load_yaml() (yaml.safe_load: the reference's bare yaml.load(fp) no longer works under PyYAML 6).
"""
import copy

N_CLASSES = 11  # airsim_loader.py:217

_COMM_DEFAULTS = dict(shared_policy=True, shared_img_encoder="unified", attention="general", sparse=False, query=True,
                      feat_squeezer=-1, feat_channel=512)


def make_config(arch, agent_num=5, img_size=512, backbones="n_segnet", query_size=32, key_size=1024, **model_overrides):
    """Build a config dict shaped like the shipped YAMLs. backbones: 'n_segnet' (the 3x3-conv benchmark pair) or
    'resnet' (resnet_encoder + simple_decoder, what every shipped YAML selects)."""
    enc, dec = {"n_segnet": ("n_segnet_encoder", "n_segnet_decoder"),
                "resnet": ("resnet_encoder", "simple_decoder")}[backbones]
    model = dict(arch=arch, agent_num=agent_num, enc_backbone=enc, dec_backbone=dec, feat_squeezer=-1,
                 feat_channel=512)
    if arch in ("MIMOcom", "MIMOcomWho", "LearnWhen2Com", "LearnWho2Com"):
        model.update(_COMM_DEFAULTS)
        model.update(query_size=query_size, key_size=key_size,
                     multiple_output=arch in ("MIMOcom", "MIMOcomWho"))
    else:
        model.update(shuffle_features="None", multiple_output=arch != "All_agents")
    model.update(model_overrides)
    return {"model": model, "data": {"dataset": "airsim", "img_rows": img_size, "img_cols": img_size},
            "training": {"batch_size": 1}}


def load_yaml(path):
    import yaml
    with open(path) as fp:
        return yaml.safe_load(fp)


def with_overrides(cfg, **model_overrides):
    c = copy.deepcopy(cfg)
    c["model"].update(model_overrides)
    return c
