/*
 * Seam 1 of INTEGRATION.md: the fast-round vote tally of one process on the device.  UNCOMPILED here (no JDK).
 *
 * Mirrors FastPaxos.handleFastRoundProposal (FastPaxos.java:125-156) — the only part of FastPaxos on the hot path; the
 * jittered classic-round timer (:94-108, :194-203) and the Paxos fallback object stay with the reference's FastPaxos, which
 * delegates its private handler to this class:
 *
 *     private void handleFastRoundProposal(final FastRoundPhase2bMessage m) { tally.handleFastRoundProposal(m); }
 *
 * Same rules, in the same order: a vote for another configuration is ignored (:126-132); a second vote of a sender is
 * ignored (:134-136); nothing is counted after a decision (:138-140); the decision needs votesReceived >= N - F AND the
 * ARRIVING vote's proposal at >= N - F, F = floor((N - 1) / 4) (:145-150).
 *
 * Proposal identity: the Java compares List<Endpoint> element by element (List.equals); the device compares the
 * order-independent 128-bit fingerprint + length of the id list (rapid_proposal_fingerprint).  Two votes that list the same
 * endpoints in DIFFERENT orders are therefore one proposal here and two in the Java — every proposer sorts its cut by the
 * ring-0 comparator before voting (MembershipService.java:346-348), so well-formed votes never differ in order only.
 */
package com.vrg.rapid;

import com.vrg.rapid.gpu.Native;
import com.vrg.rapid.pb.Endpoint;
import com.vrg.rapid.pb.FastRoundPhase2bMessage;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.function.Consumer;

final class GpuFastPaxosTally {
    private final GpuMembershipView view;
    private final long configurationId;
    private final long handle;
    private final Consumer<List<Endpoint>> onDecide;
    /** fingerprint -> the endpoint list first seen with it (what onDecide receives, like the Java's map key) */
    private final Map<Long, List<Endpoint>> firstSeen = new HashMap<>();
    private boolean decided = false;

    GpuFastPaxosTally(final GpuMembershipView view, final long configurationId, final int membershipSize,
                      final Consumer<List<Endpoint>> onDecide) {
        this.view = view;
        this.configurationId = configurationId;
        this.onDecide = onDecide;
        this.handle = Native.fpCreate(configurationId, membershipSize, membershipSize, 0);
        if (handle == 0) {
            throw new IllegalStateException(Native.lastError());
        }
    }

    /** One vote (FastPaxos.java:125-156). */
    void handleFastRoundProposal(final FastRoundPhase2bMessage proposalMessage) {
        final List<FastRoundPhase2bMessage> one = new ArrayList<>(1);
        one.add(proposalMessage);
        handleFastRoundProposals(one);
    }

    /**
     * A burst of votes in arrival order (what a busy protocol thread finds in its queue): one device call; the result is
     * the one the Java reaches by handling them one at a time — including WHICH vote decides, because the tally recovers
     * the exact arrival index at which the quorum is reached.
     */
    void handleFastRoundProposals(final List<FastRoundPhase2bMessage> votes) {
        if (decided || votes.isEmpty()) {
            return;                                                    // :138-140
        }
        final int n = votes.size();
        final int[] sender = new int[n];
        final long[] cfg = new long[n];
        final long[] h1 = new long[n];
        final long[] h2 = new long[n];
        final int[] len = new int[n];
        for (int i = 0; i < n; i++) {
            final FastRoundPhase2bMessage m = votes.get(i);
            cfg[i] = m.getConfigurationId();
            // sender membership is NOT checked by the reference (:141): an unknown sender still counts, under a joiner id
            sender[i] = view.idOf(m.getSender(), true);
            final int[] ids = new int[m.getEndpointsCount()];
            for (int j = 0; j < ids.length; j++) {
                ids[j] = view.idOf(m.getEndpoints(j), true);
            }
            final long[] fp = Native.proposalFingerprint(ids);
            h1[i] = fp[0];
            h2[i] = fp[1];
            len[i] = ids.length;
            if (cfg[i] == configurationId) {
                firstSeen.putIfAbsent(fp[0], m.getEndpointsList());
            }
        }
        final long[] out = new long[6];                                // {decided, hash, hash2, len, count, votesReceived}
        final int rc = Native.fpTally(handle, sender, cfg, h1, h2, len, out);
        if (rc != 0) {
            throw new IllegalStateException(Native.lastError());
        }
        if (out[0] != 0) {
            decided = true;
            onDecide.accept(firstSeen.get(out[1]));                    // onDecidedWrapped.accept(...), :148
        }
    }

    boolean isDecided() {
        return decided;
    }

    void shutdown() {
        Native.fpDestroy(handle);
    }
}
