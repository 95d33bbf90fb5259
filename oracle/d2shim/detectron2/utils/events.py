def get_event_storage():
    raise RuntimeError("training-only symbol; not available in the oracle shim")
