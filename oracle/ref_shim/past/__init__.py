"""Minimal stand-in for the un-vendored third-party package ``future`` (module ``past``).

TEST INFRASTRUCTURE ONLY. The reference imports ``past.utils.old_div``
(/root/reference/basepop.py:10); the package ``future`` (declared ``future>=0.15.0`` in
/root/reference/requirements.txt:5, no pinned version) is not installed in this image, so the
oracle harness puts this directory on ``sys.path`` before importing the reference.
"""
