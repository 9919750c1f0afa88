"""B200 drop-in for python/payload_sink.py: delivers received packets to a callback from a daemon
watcher thread (python/payload_sink.py:29-64)."""
import queue
import threading


class payload_sink(object):
    def __init__(self, callback=''):
        self.callback = callback
        self.sink_queue = queue.Queue()
        self._watcher = _queue_watcher_thread_mod(self.sink_queue, self)

    def set_callback(self, callback):
        self.callback = callback

    def deliver(self, payloads):
        """Push packets (bytes) as the RX chain emits them, e.g. RxResult.payloads()."""
        for p in payloads:
            self.sink_queue.put(p)

    def drain(self, timeout=5.0):
        """Block until every delivered packet has been handed to the callback."""
        self.sink_queue.join()


class _queue_watcher_thread_mod(threading.Thread):
    def __init__(self, rcvd_pktq, owner):
        threading.Thread.__init__(self)
        self.daemon = True
        self.rcvd_pktq = rcvd_pktq
        self.owner = owner
        self.keep_running = True
        self.start()

    def run(self):
        while self.keep_running:
            payload = self.rcvd_pktq.get()
            try:
                if self.owner.callback:
                    self.owner.callback(payload)
            finally:
                self.rcvd_pktq.task_done()
