"""Parity cases shared by the golden generator, the oracle tests (CPU) and the CUDA parity tests (GPU).

Every case is fully described by (config, weight seed, input seed, forward kwargs), so the inputs never need to be
stored: multiagentperception_b200.synth regenerates them bit-identically anywhere.
"""
from multiagentperception_b200 import configs

WEIGHT_SEED = 1337
INPUT_SEED = 7
RANDOM_SEED = 20   # random.seed() before every forward: the random-selection baselines draw from Python's `random`
IMG = 128     # smallest size every arch accepts: policy_net4 needs (H/32) % 4 == 0
BATCH = 2

_MO = dict(training=False, MO_flag=True)

# name -> (arch, backbones, model overrides, forward kwargs, agents in the input)
CASES = {
    "single_segnet": ("Single_agent", "n_segnet", {}, {}, 1),
    "single_resnet": ("Single_agent", "resnet", {}, {}, 1),
    "mimocom_segnet_softmax": ("MIMOcom", "n_segnet", dict(agent_num=3), dict(inference="softmax", **_MO), 3),
    "mimocom_segnet_activated": ("MIMOcom", "n_segnet", dict(agent_num=3), dict(inference="activated", **_MO), 3),
    "mimocom_segnet_argmax": ("MIMOcom", "n_segnet", dict(agent_num=3), dict(inference="argmax_test", **_MO), 3),
    "mimocom_segnet_train_sig": ("MIMOcom", "n_segnet", dict(agent_num=2), dict(training=True, MO_flag=True), 2),
    "mimocom_resnet_activated": ("MIMOcom", "resnet", dict(agent_num=6), dict(inference="activated", **_MO), 6),
    "mimocomwho_resnet_activated": ("MIMOcomWho", "resnet", dict(agent_num=4, query=False),
                                    dict(inference="activated", **_MO), 4),
    "mimocomwho_segnet_argmax": ("MIMOcomWho", "n_segnet", dict(agent_num=3), dict(inference="argmax_test", **_MO), 3),
    "when2com_resnet_activated": ("LearnWhen2Com", "resnet", dict(query_size=8),
                                  dict(training=False, inference="activated"), 5),
    "when2com_segnet_argmax": ("LearnWhen2Com", "n_segnet", dict(query_size=8),
                               dict(training=False, inference="argmax_test"), 5),
    "when2com_resnet_sparse": ("LearnWhen2Com", "resnet", dict(query_size=8, sparse=True),
                               dict(training=False, inference="softmax"), 5),
    "when2com_resnet_scaled": ("LearnWhen2Com", "resnet", dict(attention="scaled", query_size=128, key_size=128),
                               dict(training=True), 5),
    "who2com_resnet_argmax": ("LearnWho2Com", "resnet", dict(query_size=8),
                              dict(training=False, inference="argmax_test"), 5),
    "mimo_all_resnet": ("MIMO_All_agents", "resnet", dict(agent_num=3), {}, 3),
    # shipped-YAML variants added in the second session (srms_who2com.yml, *_randcom.yml) and the additive attention
    "who2com_resnet_normal_agents": ("LearnWho2Com", "resnet", dict(query_size=8, shared_img_encoder="only_normal_agents"),
                                     dict(training=False, inference="argmax_test"), 5),
    "when2com_segnet_separate_additive": ("LearnWhen2Com", "n_segnet",
                                          dict(query_size=128, key_size=128, attention="additive",
                                               shared_img_encoder=False),
                                          dict(training=False, inference="softmax"), 5),
    "mimo_all_resnet_selection": ("MIMO_All_agents", "resnet", dict(agent_num=4, shuffle_features="selection"), {}, 4),
    "mimo_all_segnet_comnet": ("MIMO_All_agents", "n_segnet", dict(agent_num=3, shuffle_features="ComNet"), {}, 3),
    "single_segnet_squeeze4": ("Single_agent", "n_segnet", dict(feat_squeezer=4), {}, 1),
    "mimo_all_resnet_squeeze2": ("MIMO_All_agents", "resnet", dict(agent_num=2, feat_squeezer=2), {}, 2),
    "all_agents_resnet_selection": ("All_agents", "resnet", dict(agent_num=5, shuffle_features="selection"), {}, 5),
    "all_agents_resnet": ("All_agents", "resnet", dict(agent_num=5), {}, 5),
}


def case_config(name):
    arch, bb, overrides, kw, n = CASES[name]
    return configs.make_config(arch, img_size=IMG, backbones=bb, **overrides), dict(kw), n


def as_tuple(out):
    return out if isinstance(out, tuple) else (out,)
