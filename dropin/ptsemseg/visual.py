"""`ptsemseg.visual`: test.py:14 does `from ptsemseg.visual import draw_bounding`, a module the reference repository
does not contain. The name is never called on the evaluation path; a no-op keeps the import
working so test.py runs unchanged."""


def draw_bounding(*args, **kwargs):
    return None
