"""TEST INFRASTRUCTURE ONLY (oracle side) -- minimal stand-in for detectron2 v0.5.

detectron2 / fvcore / pycocotools are not installable in this environment (no network), so
the reference's own Python (``/root/reference/centermask2/centermask``) is imported *unchanged*
on top of this shim to (a) validate ``oracle/restate.py`` and (b) generate the golden vectors
under ``tests/golden/``.  Only the symbols the reference imports are provided (SURVEY.md
Appendix B); training-only symbols are stubs that raise when called.  Semantics restated from
detectron2 v0.5.  Nothing under ``centermask2_b200/`` imports this.
"""
__cm2_shim__ = True
__version__ = "0.5-shim"
