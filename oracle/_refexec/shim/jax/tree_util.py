from .tree import PyTreeDef, tree_flatten, tree_leaves, tree_map, tree_structure, tree_unflatten  # noqa: F401
