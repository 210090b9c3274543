import warnings


def warn(msg, *args, category=None, stacklevel=1):
    warnings.warn(msg % args if args else msg, stacklevel=stacklevel + 1)


def deprecation(msg, *args):
    warn(msg, *args)


def error(msg, *args):
    print(msg % args if args else msg)
