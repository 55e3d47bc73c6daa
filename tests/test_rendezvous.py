"""Host rendezvous of the multi-GPU `main` (host/rendezvous.h): all-gather + barrier between forked ranks over TCP on
127.0.0.1 (CPU only; the reference gets the same two collectives from MPI, src/main.cc:8)."""
import socket

import pytest

from helpers import pkg


@pytest.mark.parametrize("world", [1, 2, 4])
def test_rendezvous_all_gather_between_forked_ranks(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    L = pkg().hostapi.lib()
    assert L.step50_rendezvous_selftest(world, port) == 0
