#include "world.h"

#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

static int env_int(const char* k, int dflt) {
  const char* v = getenv(k);
  return (v && *v) ? atoi(v) : dflt;
}

void World::from_env() {
  me = env_int("RANK", env_int("OMPI_COMM_WORLD_RANK", env_int("PMI_RANK", 0)));
  nprocs = env_int("WORLD_SIZE", env_int("OMPI_COMM_WORLD_SIZE", env_int("PMI_SIZE", 1)));
  device = env_int("LOCAL_RANK", env_int("OMPI_COMM_WORLD_LOCAL_RANK", me));
  if (nprocs < 1) nprocs = 1;
}

static bool send_all(int fd, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) {
    ssize_t k = ::send(fd, c, n, 0);
    if (k <= 0) return false;
    c += k;
    n -= (size_t)k;
  }
  return true;
}
static bool recv_all(int fd, void* p, size_t n) {
  char* c = (char*)p;
  while (n) {
    ssize_t k = ::recv(fd, c, n, 0);
    if (k <= 0) return false;
    c += k;
    n -= (size_t)k;
  }
  return true;
}

int World::bootstrap_nccl_id(unsigned char id[128], int port_offset, std::string* err) {
  if (nprocs == 1) {
    memset(id, 0, 128);
    return 0;
  }
  const char* addr = getenv("MASTER_ADDR");
  if (!addr || !*addr) addr = "127.0.0.1";
  const int port = env_int("MASTER_PORT", 29500) + port_offset;
  if (me == 0) {
    if (mmd_comm_nccl_unique_id(id) != MMD_OK) {
      if (err) *err = mmd_last_error();
      return 1;
    }
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in sa;
    memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    sa.sin_port = htons((unsigned short)port);
    if (bind(ls, (sockaddr*)&sa, sizeof sa) != 0 || listen(ls, nprocs) != 0) {
      if (err) *err = std::string("rank 0 cannot listen on port ") + std::to_string(port) + ": " + strerror(errno);
      close(ls);
      return 1;
    }
    for (int k = 1; k < nprocs; k++) {
      int fd = accept(ls, nullptr, nullptr);
      if (fd < 0 || !send_all(fd, id, 128)) {
        if (err) *err = "rank 0: sending the NCCL id failed";
        if (fd >= 0) close(fd);
        close(ls);
        return 1;
      }
      char ack;
      recv_all(fd, &ack, 1);
      close(fd);
    }
    close(ls);
    return 0;
  }
  addrinfo hints, *res = nullptr;
  memset(&hints, 0, sizeof hints);
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(addr, std::to_string(port).c_str(), &hints, &res) != 0 || !res) {
    if (err) *err = std::string("cannot resolve MASTER_ADDR ") + addr;
    return 1;
  }
  for (int attempt = 0; attempt < 600; attempt++) {  // up to ~60 s for rank 0 to come up
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
      bool ok = recv_all(fd, id, 128);
      char ack = 1;
      send_all(fd, &ack, 1);
      close(fd);
      freeaddrinfo(res);
      if (!ok && err) *err = "receiving the NCCL id failed";
      return ok ? 0 : 1;
    }
    close(fd);
    usleep(100000);
  }
  freeaddrinfo(res);
  if (err) *err = std::string("cannot reach rank 0 at ") + addr + ":" + std::to_string(port);
  return 1;
}

static void reduce(const World& w, double* v, int n, int op) {
  if (w.nprocs <= 1) return;
  if (w.reduce_cb) {
    w.reduce_cb(v, n, op, w.reduce_user);
  } else if (w.ctx) {
    if (mmd_comm_allreduce(w.ctx, v, n, op)) {
      fprintf(stderr, "ERROR: allreduce over %d ranks failed: %s\n", w.nprocs, mmd_last_error());
      exit(1);
    }
  } else {
    fprintf(stderr, "ERROR: multi-rank reduction requested but neither a device context nor a reduce callback is set\n");
    exit(1);
  }
}
void World::sum(double* v, int n) const { reduce(*this, v, n, 0); }
void World::max(double* v, int n) const { reduce(*this, v, n, 1); }
long long World::sum_ll(long long v) const {
  // counts are < 2^53: exact in a double
  double d = (double)v;
  sum(&d, 1);
  return (long long)(d + 0.5);
}
void World::barrier() const {
  double d = 0;
  sum(&d, 1);
}
