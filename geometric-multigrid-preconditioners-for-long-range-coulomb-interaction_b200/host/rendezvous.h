// Process rendezvous of the multi-GPU drop-in: what MPI_Init / MPI_Comm_rank / MPI_Allgather / MPI_Barrier give the
// reference (src/main.cc:8, src/step-50.cc:116-122), reduced to the two collectives the B200 path needs on the HOST:
// an all-gather of small blobs (the 64-byte CUDA IPC handles of gmg_dist_init) and a barrier.  All data-path
// communication is peer memory inside the kernels (csrc/dist.cuh); this is control plane only.
//
// One process per GPU.  Rank and world size come from the launcher's environment (torchrun: RANK / WORLD_SIZE /
// LOCAL_RANK / MASTER_ADDR / MASTER_PORT; Open MPI / PMI: OMPI_COMM_WORLD_RANK, PMI_RANK, ...; `main -np N file.prm`
// forks the ranks itself and sets the torchrun variables).  Rank 0 listens on MASTER_ADDR : MASTER_PORT + 29 (the
// launcher's own store keeps MASTER_PORT), the others connect; the connections stay open for the whole run.
#pragma once
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace Step50 {

class Rendezvous {
public:
  int rank = 0, world = 1, local_rank = 0;

  static int env_int(std::initializer_list<const char *> names, int fallback) {
    for (const char *n : names)
      if (const char *v = std::getenv(n))
        if (*v) return std::atoi(v);
    return fallback;
  }

  Rendezvous() {
    rank = env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0);
    world = env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1);
    local_rank = env_int({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID"}, rank);
    if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("rendezvous: inconsistent RANK / WORLD_SIZE");
  }
  Rendezvous(const Rendezvous &) = delete;
  Rendezvous &operator=(const Rendezvous &) = delete;
  ~Rendezvous() {
    for (int fd : peers)
      if (fd >= 0) ::close(fd);
    if (listen_fd >= 0) ::close(listen_fd);
  }

  // every rank contributes `bytes` bytes; out receives world * bytes, in rank order (also a barrier)
  void all_gather(const void *in, void *out, size_t bytes) {
    if (world == 1) {
      std::memcpy(out, in, bytes);
      return;
    }
    connect_all();
    std::vector<char> all(bytes * world);
    if (rank == 0) {
      std::memcpy(all.data(), in, bytes);
      for (int r = 1; r < world; ++r) recv_all(peers[r], all.data() + bytes * r, bytes);
      for (int r = 1; r < world; ++r) send_all(peers[r], all.data(), all.size());
    } else {
      send_all(peers[0], in, bytes);
      recv_all(peers[0], all.data(), all.size());
    }
    std::memcpy(out, all.data(), all.size());
  }
  void barrier() {
    char c = 0;
    std::vector<char> all(world);
    all_gather(&c, all.data(), 1);
  }

private:
  std::vector<int> peers;  // rank 0: socket per rank; others: peers[0] = the connection to rank 0
  int listen_fd = -1;
  bool connected = false;

  static void send_all(int fd, const void *p, size_t n) {
    const char *c = (const char *)p;
    while (n > 0) {
      const ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL);
      if (k <= 0) throw std::runtime_error("rendezvous: a peer process went away (send)");
      c += k;
      n -= (size_t)k;
    }
  }
  static void recv_all(int fd, void *p, size_t n) {
    char *c = (char *)p;
    while (n > 0) {
      const ssize_t k = ::recv(fd, c, n, 0);
      if (k <= 0) throw std::runtime_error("rendezvous: a peer process went away (recv)");
      c += k;
      n -= (size_t)k;
    }
  }
  void connect_all() {
    if (connected) return;
    const char *addr_env = std::getenv("MASTER_ADDR");
    const std::string addr = (addr_env && *addr_env && std::string(addr_env) != "localhost") ? addr_env : "127.0.0.1";
    const int port = env_int({"GMG_RENDEZVOUS_PORT"}, env_int({"MASTER_PORT"}, 29400) + 29);
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t)port);
    if (::inet_pton(AF_INET, addr.c_str(), &sa.sin_addr) != 1) ::inet_pton(AF_INET, "127.0.0.1", &sa.sin_addr);
    const int one = 1;
    if (rank == 0) {
      listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
      ::setsockopt(listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
      sockaddr_in any = sa;
      any.sin_addr.s_addr = htonl(INADDR_ANY);
      if (listen_fd < 0 || ::bind(listen_fd, (sockaddr *)&any, sizeof(any)) != 0 || ::listen(listen_fd, world) != 0)
        throw std::runtime_error("rendezvous: rank 0 cannot listen on port " + std::to_string(port));
      peers.assign(world, -1);
      for (int k = 1; k < world; ++k) {
        const int fd = ::accept(listen_fd, nullptr, nullptr);
        if (fd < 0) throw std::runtime_error("rendezvous: accept failed");
        ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        int32_t r = -1;
        recv_all(fd, &r, sizeof(r));
        if (r < 1 || r >= world || peers[r] >= 0) throw std::runtime_error("rendezvous: unexpected rank announced");
        peers[r] = fd;
      }
    } else {
      int fd = -1;
      const auto t0 = std::chrono::steady_clock::now();
      while (true) {
        fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (fd >= 0 && ::connect(fd, (sockaddr *)&sa, sizeof(sa)) == 0) break;
        if (fd >= 0) ::close(fd);
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
          throw std::runtime_error("rendezvous: rank 0 not reachable on port " + std::to_string(port));
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
      }
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      const int32_t r = rank;
      send_all(fd, &r, sizeof(r));
      peers.assign(1, fd);
    }
    connected = true;
  }
};

}  // namespace Step50
