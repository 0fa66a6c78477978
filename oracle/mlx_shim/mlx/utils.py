"""mlx.utils subset."""


def tree_unflatten(items):
    out = {}
    for k, v in items:
        cur = out
        parts = k.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = v
    return out


def tree_flatten(tree, prefix=""):
    flat = []
    if isinstance(tree, dict):
        for k, v in tree.items():
            flat += tree_flatten(v, f"{prefix}.{k}" if prefix else k)
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            flat += tree_flatten(v, f"{prefix}.{i}" if prefix else str(i))
    else:
        flat.append((prefix, tree))
    return flat
