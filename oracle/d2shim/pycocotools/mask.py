def __getattr__(name):
    raise RuntimeError("pycocotools is not available offline (training / evaluation only)")
