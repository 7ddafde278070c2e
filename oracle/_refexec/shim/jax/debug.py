def callback(fn, *args, ordered=False, partitioned=False, **kwargs):
    fn(*args, **kwargs)


def print(fmt, *args, **kwargs):  # noqa: A001
    __builtins__["print"](fmt.format(*args, **kwargs)) if isinstance(__builtins__, dict) else None
