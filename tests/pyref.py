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
