"""TEST INFRASTRUCTURE ONLY -- mask_head.py:10 imports pycocotools.mask at module import."""
