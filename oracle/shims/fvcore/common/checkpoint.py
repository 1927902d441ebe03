"""Test-only stand-in for fvcore.common.checkpoint.Checkpointer: only the empty-path no-op is needed."""


class Checkpointer:
    def __init__(self, model, save_dir="", *, save_to_disk=True, **checkpointables):
        self.model = model
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk
        self.checkpointables = checkpointables

    def load(self, path, checkpointables=None):
        if not path:
            return {}
        raise NotImplementedError("oracle shim: only MODEL.WEIGHTS='' (random init) is supported")
