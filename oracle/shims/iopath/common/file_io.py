"""Test-only stand-in for iopath.common.file_io (local files only; there is no network)."""
import os


class PathHandler:
    def _get_supported_prefixes(self):
        return []


class HTTPURLHandler(PathHandler):
    pass


class OneDrivePathHandler(PathHandler):
    pass


class PathManager:
    def __init__(self):
        self._handlers = []

    def register_handler(self, handler, allow_override=False):
        self._handlers.append(handler)

    def open(self, path, mode="r", **kwargs):
        return open(path, mode)

    def isfile(self, path):
        return os.path.isfile(path)

    def exists(self, path):
        return os.path.exists(path)

    def get_local_path(self, path, **kwargs):
        return path
