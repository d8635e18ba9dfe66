"""Stream- and replica-safe cache for derived device tensors (BN-folded / re-laid-out weights).

* streams: the packing kernels are enqueued on whichever CUDA stream builds an entry; a later hit on ANOTHER stream
  (HostPipeline with two compute streams, side streams) must not read the packed images before those kernels ran.  The
  entry keeps the event recorded after packing; a hit from a different stream waits for it (free once completed).
* replicas: nn.DataParallel copies a module's __dict__ shallowly, so every replica (one thread per GPU) sees the same
  cache object; entries are therefore keyed by the parameters' (data_ptr, version) — distinct per replica — kept in a
  small LRU dict and mutated under a lock."""
import threading
from collections import OrderedDict

import torch


class PackedCache:
    MAX_ENTRIES = 16

    def __init__(self):
        self._entries = OrderedDict()
        self._lock = threading.Lock()
        self._last = None

    def get(self, key, build):
        """`build()` -> value (enqueues packing kernels on the current stream)."""
        with self._lock:
            hit = self._entries.get(key)
            if hit is not None:
                self._entries.move_to_end(key)
        if hit is None:
            value = build()
            event, stream = None, None
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                s = torch.cuda.current_stream()
                event = torch.cuda.Event()
                event.record(s)
                stream = (s.device_index, s.cuda_stream)
            hit = (value, event, stream)
            with self._lock:
                self._entries[key] = hit
                while len(self._entries) > self.MAX_ENTRIES:
                    self._entries.popitem(last=False)
        else:
            value, event, stream = hit
            if event is not None and not torch.cuda.is_current_stream_capturing():
                # (while a CUDA graph is being captured the packing finished long ago: rcmvs.graph warms up first, and an
                # event query is not allowed inside a capture)
                s = torch.cuda.current_stream()
                if (s.device_index, s.cuda_stream) != stream and not event.query():
                    s.wait_event(event)
        self._last = hit[0]
        return hit[0]

    @property
    def last(self):
        """Value of the most recent get() (diagnostics only)."""
        return self._last

    def clear(self):
        with self._lock:
            self._entries.clear()
        self._last = None
