import equinox as eqx


def vmap(fn, in_axes=0, out_axes=0):
    def wrapped(*args, **kwargs):
        axes = list(in_axes) if isinstance(in_axes, (tuple, list)) else [in_axes] * len(args)
        return eqx._vmap_call(fn, args, kwargs, axes, {})

    return wrapped
