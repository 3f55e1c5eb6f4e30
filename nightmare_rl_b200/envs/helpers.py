"""The two helpers the reference's scripts import from ``envs.helpers`` -- same names, same results."""
import os


def class_to_dict(obj):
    """Recursive plain-dict view of a config object (what the reference's ``envs/helpers.py:3-18`` returns).

    Public attributes are visited in ``dir()`` order, i.e. ALPHABETICALLY, and the resulting dict keeps that order: the env
    sums its reward terms in the order of this dict, so the ordering is part of the contract."""
    if not hasattr(obj, "__dict__"):
        return obj                                           # a leaf: number, string, None, ...

    def view(value):
        return [class_to_dict(item) for item in value] if isinstance(value, list) else class_to_dict(value)

    return {name: view(getattr(obj, name)) for name in dir(obj) if not name.startswith("_")}


def _zero_padded(name, width=15):
    return name.rjust(width, "0")                            # "model_50.pt" < "model_100.pt" once both are padded


def get_load_path(root, load_run=-1, checkpoint=-1):
    """``<root>/<run>/<model file>`` to resume from (reference ``envs/helpers.py:20-42``): ``load_run == -1`` picks the
    lexicographically last run directory (they are named by date) and ignores ``exported``; ``checkpoint == -1`` the
    highest-numbered ``model_*.pt``.  An unreadable or empty ``root`` raises ``ValueError("No runs in this directory: ...")``
    even when an explicit ``load_run`` is given, as the reference does."""
    try:
        candidates = [entry for entry in os.listdir(root) if entry != "exported"]
        candidates.sort()
        latest = candidates[-1]
    except Exception as exc:
        raise ValueError("No runs in this directory: " + root) from exc
    run_dir = os.path.join(root, latest if load_run == -1 else load_run)
    if checkpoint != -1:
        return os.path.join(run_dir, "model_{}.pt".format(checkpoint))
    saved = [entry for entry in os.listdir(run_dir) if "model" in entry]
    return os.path.join(run_dir, max(saved, key=_zero_padded))
