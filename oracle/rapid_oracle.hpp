/*
 * ORACLE — test infrastructure, NOT product code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, link or call this.
 *
 * A literal C++17 restatement of the reference's Java hot path ("literal" = same data structures:
 * hash maps keyed by Endpoint{hostname bytes, port}, tree sets ordered by a memoised hash
 * comparator, per-cell sequential semantics).  Each class cites the Java file:line it follows;
 * paths are relative to /root/reference/rapid/src/main/java/com/vrg/rapid/.
 *
 *   oracle::MembershipView       <- MembershipView.java      (whole file)
 *   oracle::MultiNodeCutDetector <- MultiNodeCutDetector.java (whole file)
 *   oracle::AlertBatchHandler    <- MembershipService.java:300-354, :644-685 (batch semantics only)
 *   oracle::FastPaxosTally       <- FastPaxos.java:125-156    (fast round vote tally only)
 *
 * Pinning status: the reference cannot be built or run in this image (no JDK, no
 * jars) and its tests hold no golden vectors, so the pins are ports of the reference's own unit
 * tests (tests/test_oracle_cut_detection.py, test_oracle_membership_view.py,
 * test_oracle_fast_paxos.py).  Everything downstream of the ring hash is pinned that way; the
 * XXH64 reading of the un-vendored zero-allocation-hashing 0.8 jar is pinned only against the
 * public XXH64 vectors => "ring-hash parity unpinned".
 */
#ifndef RAPID_ORACLE_HPP
#define RAPID_ORACLE_HPP

#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "xxh64.h"

namespace oracle {

/* rapid.proto:13-17  message Endpoint { bytes hostname = 1; int32 port = 2; } */
struct Endpoint {
    std::string hostname;
    int32_t port = 0;
    bool operator==(const Endpoint& o) const { return port == o.port && hostname == o.hostname; }
    bool operator!=(const Endpoint& o) const { return !(*this == o); }
    bool operator<(const Endpoint& o) const {   // only for std::map keys in the oracle itself
        return hostname != o.hostname ? hostname < o.hostname : port < o.port;
    }
};

struct EndpointHash {
    size_t operator()(const Endpoint& e) const {
        // Any well-spread hash is faithful: Java's protobuf hashCode only drives HashMap iteration
        // order, which the hot path's results do not depend on.
        return std::hash<std::string>()(e.hostname) * 1000003u ^ (size_t)(uint32_t)e.port * 0x9E3779B1u;
    }
};

/* rapid.proto NodeId { int64 high; int64 low; } */
struct NodeId {
    int64_t high = 0, low = 0;
};

/* MembershipView.java:474-500 NodeIdComparator: signed compare of (high, low). */
struct NodeIdLess {
    bool operator()(const NodeId& a, const NodeId& b) const {
        if (a.high != b.high) return a.high < b.high;
        return a.low < b.low;
    }
};

enum EdgeStatus : int32_t { UP = 0, DOWN = 1 };   /* rapid.proto:112-115 */

struct NodeNotInRingException : std::runtime_error { using std::runtime_error::runtime_error; };
struct NodeAlreadyInRingException : std::runtime_error { using std::runtime_error::runtime_error; };
struct UUIDAlreadySeenException : std::runtime_error { using std::runtime_error::runtime_error; };

/* MembershipView.java:562-587 AddressComparator: memoised
 *   hash = xx_seed.hashBytes(hostname) * 31 + xx_seed.hashInt(port)   (wrapping int64)
 * ordered by signed Long.compare. */
class AddressComparator {
public:
    explicit AddressComparator(int seed) : seed_((uint64_t)(int64_t)seed) {}
    int64_t hashOf(const Endpoint& e) const {
        auto it = cache_.find(e);
        if (it != cache_.end()) return it->second;
        const int64_t h = computeHash(e);
        cache_.emplace(e, h);
        return h;
    }
    int64_t computeHash(const Endpoint& e) const {
        const uint64_t hb = orc_xxh64(e.hostname.data(), e.hostname.size(), seed_);
        const uint64_t hp = orc_xx_hash_int(e.port, seed_);
        return (int64_t)(hb * 31ULL + hp);
    }
    bool less(const Endpoint& a, const Endpoint& b) const { return hashOf(a) < hashOf(b); }
    void removeEndpoint(const Endpoint& e) const { cache_.erase(e); }
private:
    uint64_t seed_;
    mutable std::unordered_map<Endpoint, int64_t, EndpointHash> cache_;
};

struct RingLess {
    const AddressComparator* cmp;
    bool operator()(const Endpoint& a, const Endpoint& b) const { return cmp->less(a, b); }
};

/* MembershipView.java */
class MembershipView {
public:
    typedef std::set<Endpoint, RingLess> Ring;

    explicit MembershipView(int K) : K_(K) { init(); }                       /* :58-69 */
    MembershipView(const MembershipView&) = delete;            // rings hold pointers into cmps_
    MembershipView& operator=(const MembershipView&) = delete;

    MembershipView(int K, const std::vector<NodeId>& nodeIds,                /* :74-89 */
                   const std::vector<Endpoint>& endpoints) : K_(K) {
        init();
        // Same result as `set.addAll(endpoints)` on a TreeSet (first of two comparator-equal endpoints wins, the
        // other is silently dropped), built bottom-up: stable-sort by the memoised key, then append with an end
        // hint.  Construction is outside every timed region; only its result matters.
        for (int k = 0; k < K_; ++k) {
            std::vector<std::pair<int64_t, const Endpoint*>> order;
            order.reserve(endpoints.size());
            for (const Endpoint& e : endpoints) order.emplace_back(cmps_[k].hashOf(e), &e);
            std::stable_sort(order.begin(), order.end(),
                             [](const std::pair<int64_t, const Endpoint*>& a, const std::pair<int64_t, const Endpoint*>& b) {
                                 return a.first < b.first;
                             });
            for (const auto& kv : order) rings_[k].emplace_hint(rings_[k].end(), *kv.second);
        }
        for (const Endpoint& e : endpoints) allNodes_.insert(e);
        for (const NodeId& id : nodeIds) identifiersSeen_.insert(id);
    }

    int K() const { return K_; }

    /* :123-160 */
    void ringAdd(const Endpoint& node, const NodeId& nodeId) {
        if (isIdentifierPresent(nodeId)) throw UUIDAlreadySeenException(node.hostname);
        if (rings_[0].find(node) != rings_[0].end()) throw NodeAlreadyInRingException(node.hostname);
        std::unordered_set<Endpoint, EndpointHash> affected;
        for (int k = 0; k < K_; ++k) {
            Ring& r = rings_[k];
            r.insert(node);
            const Endpoint* subject = lower(r, node);
            if (subject) affected.insert(*subject);
        }
        allNodes_.insert(node);
        for (const Endpoint& s : affected) cachedObservers_.erase(s);
        identifiersSeen_.insert(nodeId);
        shouldUpdateConfigurationId_ = true;
    }

    /* :167-201 */
    void ringDelete(const Endpoint& node) {
        if (rings_[0].find(node) == rings_[0].end()) throw NodeNotInRingException(node.hostname);
        std::unordered_set<Endpoint, EndpointHash> affected;
        for (int k = 0; k < K_; ++k) {
            Ring& r = rings_[k];
            const Endpoint* oldSubject = lower(r, node);
            if (oldSubject) affected.insert(*oldSubject);
            r.erase(node);
            cmps_[k].removeEndpoint(node);
            cachedObservers_.erase(node);
        }
        allNodes_.erase(node);
        for (const Endpoint& s : affected) cachedObservers_.erase(s);
        shouldUpdateConfigurationId_ = true;
    }

    /* :210-224 */
    const std::vector<Endpoint>& getObserversOf(const Endpoint& node) {
        if (!allNodes_.count(node)) throw NodeNotInRingException(node.hostname);
        auto it = cachedObservers_.find(node);
        if (it == cachedObservers_.end()) it = cachedObservers_.emplace(node, computeObserversOf(node)).first;
        return it->second;
    }

    /* :267-282 */
    std::vector<Endpoint> getSubjectsOf(const Endpoint& node) const {
        if (!allNodes_.count(node)) throw NodeNotInRingException(node.hostname);
        if (rings_[0].size() <= 1) return {};
        return getPredecessorsOf(node);
    }

    /* :292-303 — predecessors, also for a node that is not (yet) a member. */
    std::vector<Endpoint> getExpectedObserversOf(const Endpoint& node) const {
        if (rings_[0].empty()) return {};
        return getPredecessorsOf(node);
    }

    bool isHostPresent(const Endpoint& e) const { return allNodes_.count(e) != 0; }          /* :330-337 */
    bool isIdentifierPresent(const NodeId& id) const { return identifiersSeen_.count(id) != 0; } /* :345-352 */

    /* isSafeToJoin :99-115 -> 0 SAFE_TO_JOIN, 1 HOSTNAME_ALREADY_IN_RING, 2 UUID_ALREADY_IN_RING */
    int isSafeToJoin(const Endpoint& node, const NodeId& id) const {
        if (allNodes_.count(node)) return 1;
        if (identifiersSeen_.count(id)) return 2;
        return 0;
    }

    /* :360-372 */
    int64_t getCurrentConfigurationId() {
        if (shouldUpdateConfigurationId_) {
            currentConfigurationId_ = configurationId(identifiersSeen_, rings_[0]);
            shouldUpdateConfigurationId_ = false;
        }
        return currentConfigurationId_;
    }

    std::vector<Endpoint> getRing(int k) const {                              /* :380-388 */
        return std::vector<Endpoint>(rings_[k].begin(), rings_[k].end());
    }

    /* :397-418 */
    std::vector<int> getRingNumbers(const Endpoint& observer, const Endpoint& subject) const {
        const std::vector<Endpoint> subjects = getSubjectsOf(observer);
        std::vector<int> out;
        int ring = 0;
        for (const Endpoint& n : subjects) {
            if (n == subject) out.push_back(ring);
            ++ring;
        }
        return out;
    }

    int getMembershipSize() const { return (int)rings_[0].size(); }           /* :425-432 */

    const AddressComparator& getRingZeroComparator() const { return cmps_[0]; } /* :468-470 */
    const AddressComparator& comparator(int k) const { return cmps_[k]; }

    /* :544-556 Configuration.getConfigurationId */
    template <class Ids, class Eps>
    static int64_t configurationId(const Ids& identifiers, const Eps& endpoints) {
        uint64_t hash = 1;
        for (const NodeId& id : identifiers) {
            hash = hash * 37 + orc_xx_hash_long(id.high, 0);
            hash = hash * 37 + orc_xx_hash_long(id.low, 0);
        }
        for (const Endpoint& e : endpoints) {
            hash = hash * 37 + orc_xxh64(e.hostname.data(), e.hostname.size(), 0);
            hash = hash * 37 + orc_xx_hash_int(e.port, 0);
        }
        return (int64_t)hash;
    }

    const std::set<NodeId, NodeIdLess>& identifiersSeen() const { return identifiersSeen_; }

private:
    void init() {
        cmps_.reserve(K_);
        for (int k = 0; k < K_; ++k) cmps_.emplace_back(k);
        rings_.reserve(K_);
        for (int k = 0; k < K_; ++k) rings_.emplace_back(RingLess{&cmps_[k]});
    }

    static const Endpoint* lower(const Ring& r, const Endpoint& node) {       // TreeSet.lower
        auto it = r.lower_bound(node);
        if (it == r.begin()) return nullptr;
        --it;
        return &*it;
    }
    static const Endpoint* higher(const Ring& r, const Endpoint& node) {      // TreeSet.higher
        auto it = r.upper_bound(node);
        if (it == r.end()) return nullptr;
        return &*it;
    }

    /* :234-257 — observers are ring SUCCESSORS */
    std::vector<Endpoint> computeObserversOf(const Endpoint& node) const {
        if (rings_[0].find(node) == rings_[0].end()) throw NodeNotInRingException(node.hostname);
        if (rings_[0].size() <= 1) return {};
        std::vector<Endpoint> out;
        for (int k = 0; k < K_; ++k) {
            const Ring& r = rings_[k];
            const Endpoint* succ = higher(r, node);
            out.push_back(succ ? *succ : *r.begin());
        }
        return out;
    }

    /* :308-322 — subjects / expected observers are ring PREDECESSORS */
    std::vector<Endpoint> getPredecessorsOf(const Endpoint& node) const {
        std::vector<Endpoint> out;
        for (int k = 0; k < K_; ++k) {
            const Ring& r = rings_[k];
            const Endpoint* pred = lower(r, node);
            out.push_back(pred ? *pred : *r.rbegin());
        }
        return out;
    }

    int K_;
    std::vector<AddressComparator> cmps_;
    std::vector<Ring> rings_;
    std::set<NodeId, NodeIdLess> identifiersSeen_;
    std::unordered_map<Endpoint, std::vector<Endpoint>, EndpointHash> cachedObservers_;
    std::unordered_set<Endpoint, EndpointHash> allNodes_;
    int64_t currentConfigurationId_ = -1;
    bool shouldUpdateConfigurationId_ = true;
};

/* rapid.proto:101-110 AlertMessage (hot-path fields only) */
struct AlertMessage {
    Endpoint edgeSrc, edgeDst;
    int32_t edgeStatus = UP;
    int64_t configurationId = 0;
    std::vector<int32_t> ringNumber;
};

/* MultiNodeCutDetector.java */
class MultiNodeCutDetector {
public:
    MultiNodeCutDetector(int K, int H, int L) : K_(K), H_(H), L_(L) {        /* :51-60 */
        if (H > K || L > H || K < 3 || L <= 0 || H <= 0)
            throw std::invalid_argument("Arguments do not satisfy K > H >= L >= 0");
    }

    int getNumProposals() const { return proposalCount_; }                    /* :62-66 */

    /* :76-82 */
    std::vector<Endpoint> aggregateForProposal(const AlertMessage& msg) {
        std::vector<Endpoint> proposals;
        for (int32_t ring : msg.ringNumber) {
            std::vector<Endpoint> r = aggregateForProposal(msg.edgeSrc, msg.edgeDst, msg.edgeStatus, ring);
            proposals.insert(proposals.end(), r.begin(), r.end());
        }
        return proposals;
    }

    /* :137-164 */
    std::vector<Endpoint> invalidateFailingEdges(MembershipView& view) {
        if (!seenLinkDownEvents_) return {};
        std::vector<Endpoint> proposalsToReturn;
        const std::vector<Endpoint> preProposalCopy(preProposal_.begin(), preProposal_.end());
        for (const Endpoint& nodeInFlux : preProposalCopy) {
            const bool present = view.isHostPresent(nodeInFlux);
            const std::vector<Endpoint> observers =
                present ? view.getObserversOf(nodeInFlux) : view.getExpectedObserversOf(nodeInFlux);
            int ringNumber = 0;
            for (const Endpoint& observer : observers) {
                if (proposal_.count(observer) || preProposal_.count(observer)) {
                    const int32_t status = present ? DOWN : UP;
                    std::vector<Endpoint> r = aggregateForProposal(observer, nodeInFlux, status, ringNumber);
                    proposalsToReturn.insert(proposalsToReturn.end(), r.begin(), r.end());
                }
                ++ringNumber;
            }
        }
        return proposalsToReturn;
    }

    void clear() {                                                            /* :169-178 */
        reportsPerHost_.clear();
        proposal_.clear();
        updatesInProgress_ = 0;
        proposalCount_ = 0;
        preProposal_.clear();
        seenLinkDownEvents_ = false;
    }

    /* test-only introspection (no Java counterpart): bitmask of reported rings for dst */
    uint32_t reportMask(const Endpoint& dst) const {
        auto it = reportsPerHost_.find(dst);
        if (it == reportsPerHost_.end()) return 0;
        uint32_t m = 0;
        for (const auto& kv : it->second) if (kv.first >= 0 && kv.first < 32) m |= 1u << kv.first;
        return m;
    }
    int updatesInProgress() const { return updatesInProgress_; }
    bool seenLinkDownEvents() const { return seenLinkDownEvents_; }

private:
    /* :84-128 */
    std::vector<Endpoint> aggregateForProposal(const Endpoint& linkSrc, const Endpoint& linkDst,
                                               int32_t edgeStatus, int ringNumber) {
        if (edgeStatus == DOWN) seenLinkDownEvents_ = true;
        std::unordered_map<int, Endpoint>& reportsForHost = reportsPerHost_[linkDst];
        if (reportsForHost.count(ringNumber)) return {};   // duplicate announcement, ignore
        reportsForHost.emplace(ringNumber, linkSrc);
        const int numReportsForHost = (int)reportsForHost.size();
        if (numReportsForHost == L_) {
            updatesInProgress_++;
            preProposal_.insert(linkDst);
        }
        if (numReportsForHost == H_) {
            preProposal_.erase(linkDst);
            proposal_.insert(linkDst);
            updatesInProgress_--;
            if (updatesInProgress_ == 0) {
                proposalCount_++;
                std::vector<Endpoint> ret(proposal_.begin(), proposal_.end());
                proposal_.clear();
                return ret;
            }
        }
        return {};
    }

    int K_, H_, L_;
    int proposalCount_ = 0;
    int updatesInProgress_ = 0;
    std::unordered_map<Endpoint, std::unordered_map<int, Endpoint>, EndpointHash> reportsPerHost_;
    std::unordered_set<Endpoint, EndpointHash> proposal_;
    std::unordered_set<Endpoint, EndpointHash> preProposal_;
    bool seenLinkDownEvents_ = false;
};

/* MembershipService.java:300-354 (batch driver) + :644-675 (filter).  One instance == the
 * protocol-thread state of one (virtual) node: its detector + announcedProposal flag.  The view is
 * shared read-only between virtual nodes (every process of a configuration holds an identical copy). */
class AlertBatchHandler {
public:
    AlertBatchHandler(MembershipView* view, int K, int H, int L) : view_(view), cd_(K, H, L) {}

    /* Returns the proposal this batch makes the node announce, sorted by the ring-0 comparator
     * (:346-348); empty if none (or if announcedProposal was already set, :318-319). */
    std::vector<Endpoint> handleBatch(const std::vector<AlertMessage>& batch) {
        const int64_t cfg = view_->getCurrentConfigurationId();
        if (announcedProposal_) return {};   // lazy stream never runs (:316-319)
        std::vector<Endpoint> proposal;      // Collectors.toSet(): a set; kept as unique vector
        std::unordered_set<Endpoint, EndpointHash> seen;
        for (const AlertMessage& msg : batch) {
            if (!filterAlertMessage(msg, cfg)) continue;
            for (const Endpoint& e : cd_.aggregateForProposal(msg))
                if (seen.insert(e).second) proposal.push_back(e);
        }
        for (const Endpoint& e : cd_.invalidateFailingEdges(*view_))
            if (seen.insert(e).second) proposal.push_back(e);
        if (proposal.empty()) return {};
        announcedProposal_ = true;
        const AddressComparator& c0 = view_->getRingZeroComparator();
        std::stable_sort(proposal.begin(), proposal.end(),
                         [&c0](const Endpoint& a, const Endpoint& b) { return c0.less(a, b); });
        return proposal;
    }

    /* :644-675 */
    bool filterAlertMessage(const AlertMessage& m, int64_t currentConfigurationId) const {
        if (currentConfigurationId != m.configurationId) return false;
        if (m.edgeStatus == UP && view_->isHostPresent(m.edgeDst)) return false;
        if (m.edgeStatus == DOWN && !view_->isHostPresent(m.edgeDst)) return false;
        return true;
    }

    /* decideViewChange :424-426 resets */
    void reset() { cd_.clear(); announcedProposal_ = false; }

    bool announcedProposal() const { return announcedProposal_; }
    MultiNodeCutDetector& detector() { return cd_; }

private:
    MembershipView* view_;
    MultiNodeCutDetector cd_;
    bool announcedProposal_ = false;
};

/* FastPaxos.java:125-156 handleFastRoundProposal — the fast-round tally of one node. */
class FastPaxosTally {
public:
    FastPaxosTally(int64_t configurationId, int membershipSize)
        : configurationId_(configurationId), membershipSize_(membershipSize) {}

    /* returns true iff THIS vote triggered the decision */
    bool handleFastRoundProposal(const Endpoint& sender, int64_t configurationId,
                                 const std::vector<Endpoint>& endpoints) {
        if (configurationId != configurationId_) return false;                 /* :126 */
        if (votesReceived_.count(sender)) return false;                         /* :134 */
        if (decided_) return false;                                             /* :138 */
        votesReceived_.insert(sender);                                          /* :141 */
        const int count = ++votesPerProposal_[endpoints];                       /* :142-144 */
        const int F = (int)((membershipSize_ - 1) / 4);   // floor((N-1)/4.0), exact for N >= 1  (:145)
        if ((int)votesReceived_.size() >= membershipSize_ - F) {                /* :146 */
            if (count >= membershipSize_ - F) {                                 /* :147 */
                decided_ = true;
                decision_ = endpoints;
                return true;
            }
        }
        return false;
    }

    bool decided() const { return decided_; }
    const std::vector<Endpoint>& decision() const { return decision_; }
    int votesReceived() const { return (int)votesReceived_.size(); }
    int votesFor(const std::vector<Endpoint>& p) const {
        auto it = votesPerProposal_.find(p);
        return it == votesPerProposal_.end() ? 0 : it->second;
    }

private:
    int64_t configurationId_;
    int membershipSize_;
    // Java: HashMap<List<Endpoint>, AtomicInteger>; List.hashCode/equals are O(#cut) per vote.
    // std::map with lexicographic compare has the same O(#cut)-per-probe cost driver.
    std::map<std::vector<Endpoint>, int> votesPerProposal_;
    std::unordered_set<Endpoint, EndpointHash> votesReceived_;
    bool decided_ = false;
    std::vector<Endpoint> decision_;
};

}  // namespace oracle

#endif
