"""Helper functions with the reference's behaviour (``envs/helpers.py``)."""
import os


def class_to_dict(obj):
    """Plain-dict view of a config object (reference ``envs/helpers.py:3-18``).

    Keys come from ``dir(obj)`` and are therefore ALPHABETICAL; the env relies on that for the
    order in which reward terms are summed."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        value = getattr(obj, key)
        out[key] = [class_to_dict(v) for v in value] if isinstance(value, list) else class_to_dict(value)
    return out


def get_load_path(root, load_run=-1, checkpoint=-1):
    """Newest run directory / highest-numbered ``model_*.pt`` (reference ``envs/helpers.py:20-42``)."""
    try:
        runs = sorted(r for r in os.listdir(root) if r != "exported")
        newest = os.path.join(root, runs[-1])
    except Exception as exc:
        raise ValueError("No runs in this directory: " + root) from exc
    run_dir = newest if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = sorted((f for f in os.listdir(run_dir) if "model" in f), key=lambda m: "{0:0>15}".format(m))
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(run_dir, model)
