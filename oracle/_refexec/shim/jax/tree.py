import equinox as eqx


def map(f, tree, *rest, is_leaf=None):  # noqa: A001
    return eqx.tree_map(f, tree, *rest, is_leaf=is_leaf)


def leaves(tree, is_leaf=None):
    return eqx.tree_leaves(tree, is_leaf=is_leaf)


def structure(tree, is_leaf=None):
    return eqx.tree_structure(tree, is_leaf=is_leaf)


def flatten(tree, is_leaf=None):
    return eqx.tree_flatten(tree, is_leaf=is_leaf)


def unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


tree_map, tree_leaves, tree_structure, tree_flatten, tree_unflatten = map, leaves, structure, flatten, unflatten
PyTreeDef = eqx.TreeDef
