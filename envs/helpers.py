from nightmare_rl_b200.envs.helpers import class_to_dict, get_load_path  # noqa: F401
