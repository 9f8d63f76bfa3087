"""Drop-in for ptsemseg.models: get_model(model_dict, n_classes, version=None) -> nn.Module.

Mirrors ptsemseg/models/__init__.py:8-101 — the whole YAML dict goes in, the arch name selects the class and the
keyword plumbing. Optional extra key (ignored by the reference, absent from its YAMLs):
  model.precision: 'bf16' | 'fp16' | 'bf16x3'   (default: env W2C_PRECISION, else 'bf16')
"""
from .agents import (All_agents, LearnWhen2Com, LearnWho2Com, MIMO_All_agents, MIMOcom, MIMOcomWho, Single_agent)

_REGISTRY = {
    "Single_agent": Single_agent,
    "All_agents": All_agents,
    "MIMO_All_agents": MIMO_All_agents,
    "LearnWho2Com": LearnWho2Com,
    "LearnWhen2Com": LearnWhen2Com,
    "MIMOcom": MIMOcom,
    "MIMOcomWho": MIMOcomWho,
}

_COMM_ARCHS = ("LearnWho2Com", "LearnWhen2Com", "MIMOcom", "MIMOcomWho")


def _get_model_instance(name):
    try:
        return _REGISTRY[name]
    except KeyError:
        # the reference does `raise ("Model {} not available")`, a TypeError in py3; same text, proper type
        raise ValueError("Model {} not available".format(name))


def get_model(model_dict, n_classes, version=None):
    m = model_dict["model"]
    name = m["arch"]
    cls = _get_model_instance(name)
    common = dict(n_classes=n_classes, in_channels=3, enc_backbone=m["enc_backbone"], dec_backbone=m["dec_backbone"])
    if name == "Single_agent":
        model = cls(feat_squeezer=m["feat_squeezer"], feat_channel=m["feat_channel"], **common)
    elif name in ("All_agents", "MIMO_All_agents"):
        model = cls(aux_agent_num=m["agent_num"], shuffle_flag=m["shuffle_features"],
                    feat_squeezer=m["feat_squeezer"], feat_channel=m["feat_channel"], **common)
    elif name in _COMM_ARCHS:
        kw = dict(attention=m["attention"], has_query=m["query"], sparse=m["sparse"],
                  shared_img_encoder=m["shared_img_encoder"], image_size=model_dict["data"]["img_rows"],
                  query_size=m["query_size"], key_size=m["key_size"], **common)
        if name in ("MIMOcom", "MIMOcomWho"):
            kw["agent_num"] = m["agent_num"]
        else:
            kw["aux_agent_num"] = m["agent_num"]
        model = cls(**kw)
    else:  # pragma: no cover - registry and branches cover the same names
        model = cls(**common)
    if m.get("precision"):
        model.set_precision(m["precision"])
    return model


__all__ = ["get_model"] + sorted(_REGISTRY)
