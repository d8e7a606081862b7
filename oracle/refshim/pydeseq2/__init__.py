"""Import shim (TEST INFRASTRUCTURE, container-only).

Makes the read-only reference checkout at /root/reference importable without
installing it: the real ``pydeseq2/__init__.py`` calls
``importlib.metadata.version("pydeseq2")`` which fails for an un-installed tree.
Only used by ``oracle/make_golden.py`` and the ``refcheck`` tests that run in the
build container; nothing on the product path imports it and it is never present on
the GPU box (``/root/reference`` does not exist there).
"""
__path__ = ["/root/reference/pydeseq2"]
__file__ = "/root/reference/pydeseq2/__init__.py"
__version__ = "0.5.3"
