import os
class PathManager:
    @staticmethod
    def ls(path): return os.listdir(path)
    @staticmethod
    def exists(path): return os.path.exists(path)
    @staticmethod
    def open(path, *a, **k): return open(path, *a, **k)
