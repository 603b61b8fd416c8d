"""Stand-in for `quick_queue` (absent offline): only the reference's disk datasets use it."""
import queue


class QQueue(queue.Queue):
    def __init__(self, maxsize=0, size_bucket_list=1):
        super().__init__(maxsize)

    def put_bucket(self, x):
        self.put(x)

    def get_bucket(self):
        return self.get()
