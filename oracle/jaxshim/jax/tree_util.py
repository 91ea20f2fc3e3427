"""``jax.tree_util`` stand-in (TEST INFRASTRUCTURE ONLY)."""


def tree_map(f, tree, *rest):
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*(tree_map(f, t, *(r[i] for r in rest)) for i, t in enumerate(tree)))
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, t, *(r[i] for r in rest)) for i, t in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *(r[k] for r in rest)) for k, v in tree.items()}
    return f(tree, *rest)
