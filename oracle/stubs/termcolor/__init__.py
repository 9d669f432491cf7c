"""Minimal stand-in for the `termcolor` package (absent from this image).

Test infrastructure only: lets the upstream reference be imported from
/root/reference when generating golden vectors (tests/golden/make_golden.py).
"""


def cprint(*args, **kwargs):
    text = args[0] if args else ""
    print(text, **{k: kwargs[k] for k in ("end", "file", "flush") if k in kwargs})


def colored(text, *args, **kwargs):
    return text
