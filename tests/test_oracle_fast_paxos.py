"""Port of FastPaxosWithoutFallbackTests (rapid/src/test/java/com/vrg/rapid/FastPaxosWithoutFallbackTests.java)
parameter tables against oracle::FastPaxosTally — the decision happens exactly at the quorum-th
identical vote; senders need not be members (addrForBase uses ports 0.. while the view uses 1234..)."""
import pytest

NO_CONFLICTS = [(6, 5), (48, 37), (50, 38), (100, 76), (102, 77), (5, 4), (51, 39), (49, 37), (99, 75), (101, 76)]
CONFLICTS = [
    (6, 5, 1, True), (48, 37, 1, True), (50, 38, 1, True), (100, 76, 1, True), (102, 77, 1, True),
    (48, 37, 11, True), (50, 38, 12, True), (100, 76, 24, True), (102, 77, 25, True),
    (6, 5, 2, False), (48, 37, 14, False), (50, 38, 13, False), (100, 76, 25, False), (102, 77, 26, False),
]
CFG = 0x1234


@pytest.mark.parametrize("N,quorum", NO_CONFLICTS)
def test_fast_quorum_no_conflicts(orc, N, quorum):           # :62-90
    u = orc.Universe()
    fp = orc.FastPaxosTally(u, CFG, N)
    proposal = [u.add("127.0.0.1", 1235)]
    for i in range(quorum - 1):
        assert not fp.handleFastRoundProposal(u.add("127.0.0.1", i), CFG, proposal)
        assert not fp.decided()
    assert fp.handleFastRoundProposal(u.add("127.0.0.1", quorum - 1), CFG, proposal)
    assert fp.decided() and fp.decision() == proposal
    assert quorum == N - (N - 1) // 4


@pytest.mark.parametrize("N,quorum,conflicts,change", CONFLICTS)
def test_fast_quorum_with_conflicts(orc, N, quorum, conflicts, change):   # :97-148
    u = orc.Universe()
    fp = orc.FastPaxosTally(u, CFG, N)
    proposal = [u.add("127.0.0.1", 1235)]
    conflict = [u.add("127.0.0.1", 1236)]
    for i in range(conflicts):
        fp.handleFastRoundProposal(u.add("127.0.0.1", i), CFG, conflict)
        assert not fp.decided()
    non_conflict = min(conflicts + quorum - 1, N - 1)
    for i in range(conflicts, non_conflict):
        fp.handleFastRoundProposal(u.add("127.0.0.1", i), CFG, proposal)
        assert not fp.decided()
    fp.handleFastRoundProposal(u.add("127.0.0.1", non_conflict), CFG, proposal)
    assert fp.decided() == change


def test_filters(orc):                                        # FastPaxos.java:126-140
    u = orc.Universe()
    fp = orc.FastPaxosTally(u, CFG, 5)
    p = [u.add("h", 1)]
    assert not fp.handleFastRoundProposal(u.add("s", 0), CFG + 1, p)   # wrong configuration: ignored
    assert fp.votesReceived() == 0
    for i in range(3):
        fp.handleFastRoundProposal(u.add("s", i), CFG, p)
    fp.handleFastRoundProposal(u.add("s", 0), CFG, p)                  # duplicate sender: ignored
    assert fp.votesReceived() == 3 and not fp.decided()
    assert fp.handleFastRoundProposal(u.add("s", 3), CFG, p)           # 4th distinct sender decides (Q=4)
    assert not fp.handleFastRoundProposal(u.add("s", 4), CFG, p)       # after decision: ignored entirely
    assert fp.votesReceived() == 4
