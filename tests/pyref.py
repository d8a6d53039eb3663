"""A SECOND, independent restatement of the reference's hot-path rules, in plain Python (test infrastructure, small cases only).

The C++ oracle (oracle/rapid_oracle.hpp) is what every GPU parity test compares against; it is pinned by ports of the reference's
own unit tests.  This module restates the same three pieces again, straight from the Java, with Python's own containers, so that
tests/test_oracle_vs_python_restatement.py can run random streams through BOTH readings and demand identical observables:

  PyCutDetector    MultiNodeCutDetector.java:51-178
  PyBatchHandler   MembershipService.java:300-354 (batch driver) + :644-675 (filter)
  PyFastPaxos      FastPaxos.java:125-156

The membership view (ring order, observers) is taken from the oracle's view object: what is cross-checked here is the detector /
handler / tally logic, not the ring hash."""

UP, DOWN = 0, 1
K_MIN = 3


class PyCutDetector:
    def __init__(self, K, H, L):                                   # :51-60
        if H > K or L > H or K < K_MIN or L <= 0 or H <= 0:
            raise ValueError("Arguments do not satisfy K > H >= L >= 0")
        self.K, self.H, self.L = K, H, L
        self.clear()

    def clear(self):                                               # :169-178
        self.reportsPerHost = {}
        self.proposal = set()
        self.preProposal = set()
        self.updatesInProgress = 0
        self.proposalCount = 0
        self.seenLinkDownEvents = False

    def getNumProposals(self):                                     # :62-66
        return self.proposalCount

    def aggregateForProposal(self, src, dst, status, rings):       # :76-82
        out = []
        for r in rings:
            out += self._aggregate(src, dst, status, r)
        return out

    def _aggregate(self, linkSrc, linkDst, edgeStatus, ringNumber):   # :84-128
        assert ringNumber <= self.K
        if edgeStatus == DOWN:
            self.seenLinkDownEvents = True
        reportsForHost = self.reportsPerHost.setdefault(linkDst, {})
        if ringNumber in reportsForHost:
            return []
        reportsForHost[ringNumber] = linkSrc
        n = len(reportsForHost)
        if n == self.L:
            self.updatesInProgress += 1
            self.preProposal.add(linkDst)
        if n == self.H:
            self.preProposal.discard(linkDst)                      # (HashSet.remove: no-op if absent, i.e. when L > H never held)
            self.proposal.add(linkDst)
            self.updatesInProgress -= 1
            if self.updatesInProgress == 0:
                self.proposalCount += 1
                ret = list(self.proposal)
                self.proposal.clear()
                return ret
        return []

    def invalidateFailingEdges(self, view):                        # :137-164
        if not self.seenLinkDownEvents:
            return []
        out = []
        for nodeInFlux in list(self.preProposal):
            present = view.isHostPresent(nodeInFlux)
            observers = view.getObserversOf(nodeInFlux) if present else view.getExpectedObserversOf(nodeInFlux)
            for ringNumber, observer in enumerate(observers):
                if observer in self.proposal or observer in self.preProposal:
                    out += self._aggregate(observer, nodeInFlux, DOWN if present else UP, ringNumber)
        return out

    def reportMask(self, tag):
        m = 0
        for r in self.reportsPerHost.get(tag, {}):
            m |= 1 << r
        return m


class PyBatchHandler:
    """one process's handleMessage(BatchedAlertMessage); msgs are (src, dst, status, cfg, [rings])"""

    def __init__(self, view, K, H, L):
        self.view = view
        self.cd = PyCutDetector(K, H, L)
        self.announcedProposal = False
        self.joiners_seen = []                                     # extractJoinerUuidAndMetadata: which UP alerts were looked at

    def _filter(self, msg, cfg_now):                               # :644-675
        _, dst, status, cfg, _ = msg
        if cfg_now != cfg:
            return False
        if status == UP and self.view.isHostPresent(dst):
            return False
        if status == DOWN and not self.view.isHostPresent(dst):
            return False
        return True

    def handleBatch(self, msgs):                                   # :300-354
        cfg_now = self.view.getCurrentConfigurationId()
        if self.announcedProposal:                                 # the lazy stream never runs
            return set()
        proposal = set()
        for m in msgs:
            if not self._filter(m, cfg_now):
                continue
            if m[2] == UP:
                self.joiners_seen.append(m[1])
            proposal.update(self.cd.aggregateForProposal(m[0], m[1], m[2], m[4]))
        proposal.update(self.cd.invalidateFailingEdges(self.view))
        if proposal:
            self.announcedProposal = True
        return proposal

    def reset(self):                                               # decideViewChange :425-426
        self.cd.clear()
        self.announcedProposal = False


class PyFastPaxos:
    def __init__(self, configurationId, membershipSize):
        self.configurationId, self.membershipSize = configurationId, membershipSize
        self.votesReceived = set()
        self.votesPerProposal = {}
        self.decided = False
        self.decision = None

    def handleFastRoundProposal(self, sender, cfg, endpoints):     # :125-156; True iff THIS vote decided
        if cfg != self.configurationId:
            return False
        if sender in self.votesReceived:
            return False
        if self.decided:
            return False
        self.votesReceived.add(sender)
        key = tuple(endpoints)                                     # List<Endpoint>.equals: order matters
        count = self.votesPerProposal[key] = self.votesPerProposal.get(key, 0) + 1
        F = (self.membershipSize - 1) // 4                         # floor((N - 1) / 4.0)
        if len(self.votesReceived) >= self.membershipSize - F and count >= self.membershipSize - F:
            self.decided, self.decision = True, list(endpoints)
            return True
        return False


# ----------------------------------------------------------------------------------------------------------------------------
# MembershipView.java restated a second time: python-xxhash (an XXH64 implementation that shares no code with oracle/xxh64.h),
# sorted Python lists for the TreeSets.  Same READING of zero-allocation-hashing as the oracle (hashInt / hashLong = XXH64 of the
# little-endian 4 / 8 bytes, hashBytes = XXH64 of the hostname bytes) — that reading stays unpinned; what this
# catches is an implementation slip on either side.
# ----------------------------------------------------------------------------------------------------------------------------
import bisect
import struct

M64 = (1 << 64) - 1


def _s64(x):
    x &= M64
    return x - (1 << 64) if x >> 63 else x


class PyMembershipView:
    def __init__(self, K, endpoints=(), node_ids=()):
        """endpoints: list of (tag, hostname bytes, port); node_ids: list of (high, low) signed 64-bit"""
        import xxhash
        self._xx = xxhash.xxh64_intdigest
        self.K = K
        self.ep = {}                                               # tag -> (hostname, port)
        self.rings = [[] for _ in range(K)]                        # sorted [(signed key, tag)]
        self.keys = {}                                             # (k, tag) -> key   (AddressComparator.hashCache)
        self.allNodes = set()
        self.identifiersSeen = set()
        self.cachedObservers = {}                                  # :52 — restated WITH its invalidation rule (see ringAdd)
        for tag, host, port in endpoints:
            self.ep[tag] = (host, port)
            self.allNodes.add(tag)
            for k in range(K):
                bisect.insort(self.rings[k], (self._key(k, tag), tag))
        self.identifiersSeen.update(node_ids)

    def know(self, tag, host, port):
        self.ep[tag] = (host, port)

    def _key(self, k, tag):                                        # AddressComparator.computeHash :577-582, seed = ring index
        if (k, tag) not in self.keys:
            host, port = self.ep[tag]
            h = self._xx(host, seed=k) * 31 + self._xx(struct.pack("<i", port), seed=k)
            self.keys[(k, tag)] = _s64(h)
        return self.keys[(k, tag)]

    def ringAdd(self, tag, node_id):                               # :123-160
        if node_id in self.identifiersSeen:
            raise KeyError("UUIDAlreadySeen")
        if tag in self.allNodes:
            raise ValueError("NodeAlreadyInRing")
        affected = set()
        for k in range(self.K):
            bisect.insort(self.rings[k], (self._key(k, tag), tag))
            low = self._lower_no_wrap(k, tag)                      # endpoints.lower(node): null at the front of the ring — the LAST
            if low is not None:                                    # node, whose successor wraps around to the new node, keeps its
                affected.add(low)                                  # cached list (the reference's behaviour, restated as it is)
        self.allNodes.add(tag)
        for t in affected:
            self.cachedObservers.pop(t, None)
        self.identifiersSeen.add(node_id)

    def ringDelete(self, tag):                                     # :167-201
        if tag not in self.allNodes:
            raise LookupError("NodeNotInRing")
        affected = set()
        for k in range(self.K):
            low = self._lower_no_wrap(k, tag)
            if low is not None:
                affected.add(low)
            self.rings[k].remove((self._key(k, tag), tag))
            self.cachedObservers.pop(tag, None)
        self.allNodes.discard(tag)
        for t in affected:
            self.cachedObservers.pop(t, None)

    def _lower_no_wrap(self, k, tag):
        ring = self.rings[k]
        i = bisect.bisect_left(ring, (self._key(k, tag), tag))
        return ring[i - 1][1] if i > 0 else None

    def _neighbour(self, k, tag, step):
        ring = self.rings[k]
        key = (self._key(k, tag), tag)
        if step > 0:                                               # higher(node), else first()
            i = bisect.bisect_right(ring, key)
            return ring[i][1] if i < len(ring) else ring[0][1]
        i = bisect.bisect_left(ring, key)                          # lower(node), else last()
        return ring[i - 1][1] if i > 0 else ring[-1][1]

    def getObserversOf(self, tag):                                 # :210-257
        if tag not in self.allNodes:
            raise LookupError("NodeNotInRing")
        if tag not in self.cachedObservers:
            self.cachedObservers[tag] = self.computeObserversOf(tag)
        return self.cachedObservers[tag]

    def computeObserversOf(self, tag):                             # :234-257
        if len(self.rings[0]) <= 1:
            return []
        return [self._neighbour(k, tag, +1) for k in range(self.K)]

    def getSubjectsOf(self, tag):                                  # :267-282
        if tag not in self.allNodes:
            raise LookupError("NodeNotInRing")
        if len(self.rings[0]) <= 1:
            return []
        return [self._neighbour(k, tag, -1) for k in range(self.K)]

    def getExpectedObserversOf(self, tag):                         # :292-303 (predecessors — also for a node that is not a member)
        if not self.rings[0]:
            return []
        return [self._neighbour(k, tag, -1) for k in range(self.K)]

    def getRingNumbers(self, observer, subject):                   # :397-418
        return [r for r, s in enumerate(self.getSubjectsOf(observer)) if s == subject]

    def getRing(self, k):
        return [t for _, t in self.rings[k]]

    def getMembershipSize(self):
        return len(self.rings[0])

    def isHostPresent(self, tag):
        return tag in self.allNodes

    def getCurrentConfigurationId(self):                           # Configuration.getConfigurationId :544-556
        h = 1
        for high, low in sorted(self.identifiersSeen):             # NodeIdComparator :474-500: signed (high, low)
            h = (h * 37 + self._xx(struct.pack("<q", high), seed=0)) & M64
            h = (h * 37 + self._xx(struct.pack("<q", low), seed=0)) & M64
        for _, tag in self.rings[0]:
            host, port = self.ep[tag]
            h = (h * 37 + self._xx(host, seed=0)) & M64
            h = (h * 37 + self._xx(struct.pack("<i", port), seed=0)) & M64
        return _s64(h)


# ----------------------------------------------------------------------------------------------------------------------------
# Paxos.selectProposalUsingCoordinatorRule (Paxos.java:271-328), restated a second time.  msgs: [{'vrnd': (round, nodeIndex),
# 'vval': [tags]}] in arrival order.
# ----------------------------------------------------------------------------------------------------------------------------
def coordinator_rule(N, msgs):
    if not msgs:
        raise ValueError("phase1bMessages was empty")
    maxVrnd = max(m["vrnd"] for m in msgs)                         # compareRanks :333-339: (round, nodeIndex) lexicographic
    collected = [tuple(m["vval"]) for m in msgs if m["vrnd"] == maxVrnd and len(m["vval"]) > 0]
    chosen = None
    if len(set(collected)) == 1:
        chosen = collected[0]
    elif len(collected) > 1:
        counters = {}
        for value in collected:
            count = counters.get(value, 0)
            if count + 1 > N // 4:
                chosen = value
                break
            counters[value] = count + 1
    if chosen is None:
        chosen = next((tuple(m["vval"]) for m in msgs if len(m["vval"]) > 0), ())
    return list(chosen)


# ----------------------------------------------------------------------------------------------------------------------------
# PingPongFailureDetector.java:38-121 + one process's detectors (MembershipService.java:697-707) + the AlertMessage of a
# notification (:472-495), restated a second time.  A probe's outcome comes from the scenario: 'ok' | 'fail' | 'bootstrapping'.
# ----------------------------------------------------------------------------------------------------------------------------
FAILURE_THRESHOLD, BOOTSTRAP_COUNT_THRESHOLD = 10, 30


class PyPingPong:
    def __init__(self):
        self.failureCount = 0
        self.bootstrapResponseCount = 0
        self.notified = False

    def run(self, probe):
        """-> True iff the notifier ran in this interval; `probe()` is only called when a probe is sent"""
        if self.failureCount >= FAILURE_THRESHOLD and not self.notified:
            self.notified = True
            return True
        outcome = probe()
        if outcome == "fail":
            self.failureCount += 1
        elif outcome == "bootstrapping":
            self.bootstrapResponseCount += 1
            if self.bootstrapResponseCount > BOOTSTRAP_COUNT_THRESHOLD:
                self.failureCount += 1
        return False


class PyFdCluster:
    """every member's detectors, one per entry of getSubjectsOf(member) (duplicates included), ticked in member order"""
    CRASHED, INGRESS_BLOCKED, EGRESS_BLOCKED, BOOTSTRAPPING = 1, 2, 4, 8

    def __init__(self, view, members):
        self.view, self.members = view, list(members)
        self.subjects = {m: view.getSubjectsOf(m) for m in self.members}
        self.fds = {m: [PyPingPong() for _ in self.subjects[m]] for m in self.members}

    def tick(self, flags, edge_fail=None):
        """flags[tag]: scenario bits; edge_fail[(member, detector index)]: that probe fails.  -> [(observer, subject, [rings])]"""
        alerts = []
        for m in self.members:
            if flags[m] & self.CRASHED:
                continue                                           # a crashed process runs nothing
            for j, (fd, s) in enumerate(zip(self.fds[m], self.subjects[m])):
                def probe():
                    if edge_fail and edge_fail.get((m, j)):
                        return "fail"
                    if (flags[m] & self.EGRESS_BLOCKED) or (flags[s] & (self.CRASHED | self.INGRESS_BLOCKED)):
                        return "fail"
                    if flags[s] & self.BOOTSTRAPPING:
                        return "bootstrapping"
                    return "ok"
                if fd.run(probe):
                    alerts.append((m, s, self.view.getRingNumbers(m, s)))
        return alerts


# ----------------------------------------------------------------------------------------------------------------------------
# Paxos.java:73-257 restated a second time (one object per node; handlers RETURN the message the Java hands to its broadcaster /
# client, like the oracle's binding).  Ranks are (round, nodeIndex) tuples; my_hash stands in for myAddr.hashCode() (:102).
# ----------------------------------------------------------------------------------------------------------------------------
class PyPaxos:
    def __init__(self, my_tag, my_hash, configurationId, N):
        self.me, self.my_hash, self.cfg, self.N = my_tag, my_hash, configurationId, N
        self.rnd = self.vrnd = self.crnd = (0, 0)
        self.vval, self.cval = [], []
        self.phase1bMessages = []
        self.acceptResponses = {}
        self.decided, self.decision = False, None

    def startPhase1a(self, round_):                                # :98-113
        if self.crnd[0] > round_:
            return None
        self.crnd = (round_, self.my_hash)
        return {"sender": self.me, "cfg": self.cfg, "rank": self.crnd}

    def handlePhase1aMessage(self, m):                             # :120-151
        if m["cfg"] != self.cfg:
            return None
        if self.rnd < tuple(m["rank"]):
            self.rnd = tuple(m["rank"])
        else:
            return None
        return {"sender": self.me, "cfg": self.cfg, "rnd": self.rnd, "vrnd": self.vrnd, "vval": list(self.vval)}

    def handlePhase1bMessage(self, m):                             # :159-191
        if m["cfg"] != self.cfg:
            return None
        if self.crnd != tuple(m["rnd"]):
            return None
        self.phase1bMessages.append({"vrnd": tuple(m["vrnd"]), "vval": list(m["vval"])})
        if len(self.phase1bMessages) > self.N // 2:
            chosen = coordinator_rule(self.N, self.phase1bMessages)
            if self.crnd == tuple(m["rnd"]) and not self.cval and chosen:
                self.cval = chosen
                return {"sender": self.me, "cfg": self.cfg, "rnd": self.crnd, "vval": list(chosen)}
        return None

    def handlePhase2aMessage(self, m):                             # :198-216
        if m["cfg"] != self.cfg:
            return None
        r = tuple(m["rnd"])
        if self.rnd <= r and self.vrnd != r:
            self.rnd = self.vrnd = r
            self.vval = list(m["vval"])
            return {"sender": self.me, "cfg": self.cfg, "rnd": r, "endpoints": list(self.vval)}
        return None

    def handlePhase2bMessage(self, m):                             # :223-236; True iff this message made the node decide
        if m["cfg"] != self.cfg:
            return False
        in_rnd = self.acceptResponses.setdefault(tuple(m["rnd"]), {})
        in_rnd[m["sender"]] = m
        if len(in_rnd) > self.N // 2 and not self.decided:
            self.decision = list(m["endpoints"])
            self.decided = True
            return True
        return False

    def registerFastRoundVote(self, vote):                         # :244-257
        if self.rnd[0] > 1:
            return
        self.rnd = (1, 1)
        self.vrnd = self.rnd
        self.vval = list(vote)
