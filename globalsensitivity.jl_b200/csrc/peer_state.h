// peer_state.h -- host-side state of a handle's peer group (peer.cu owns it; gsa_api.cu reads it to fuse the exchange into
// the tail of the K2+K3 kernel).
#pragma once
#include <vector>

namespace gsa {

struct PeerState {
    int rank = -1, world = 0;
    size_t max_doubles = 0;
    void* mailbox = nullptr;                 // this rank's: [2][world_max][max_doubles] doubles, then flags [2][world_max] u64
    std::vector<void*> peer;                 // mapped mailboxes, peer[rank] = own
    void** peer_dev = nullptr;               // device copy of `peer`
    unsigned long long epoch = 0, ctas_launched = 0;
    unsigned long long* counter_dev = nullptr;
    int* status_host = nullptr;              // pinned + mapped: 0 ok, 1 timeout
    int* status_dev = nullptr;
};

// fills the exchange arguments of the NEXT epoch (advances ps->epoch); data/host_out/count/nstat are the caller's
struct PeerArgs;
bool peer_connected(const gsa_handle* h);   // a group of more than one rank is connected
int peer_next_exchange(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out_mapped_dev, PeerArgs* out);

}  // namespace gsa
