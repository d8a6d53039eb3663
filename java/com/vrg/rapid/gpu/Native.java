/*
 * JNI veneer over librapid_b200.so (include/rapid_b200.h).  UNCOMPILED in this repository's build image (no JDK);
 * it is the binding a Rapid maintainer adds: one static native per C entry point, handles as long, arrays as direct
 * ByteBuffers / primitive arrays.  Every method returns the C status code (0 ok, <0 error); lastError() is the message.
 */
package com.vrg.rapid.gpu;

import java.nio.ByteBuffer;

public final class Native {
    static {
        System.loadLibrary("rapid_jni");     // java/jni/rapid_jni.c, linked against librapid_b200.so
    }

    private Native() {
    }

    public static native String lastError();

    // ---- MembershipView (com.vrg.rapid.MembershipView) ----
    /** rapid_view_create: hostBytes = hostnames concatenated, hostOff[n+1]; returns handle or 0 (see lastError). */
    public static native long viewCreate(int k, long n, byte[] hostBytes, int[] hostOff, int[] port, int device);
    public static native int viewDestroy(long view);
    public static native int viewRing(long view, int ring, int[] outIds);
    public static native int viewObservers(long view, int node, int[] outK);          // returns count or <0
    public static native int viewSubjects(long view, int node, int[] outK);
    public static native int viewExpectedObservers(long view, byte[] host, int port, int[] outK);
    public static native int viewRingNumbers(long view, int observer, int subject);   // bitmask or <0
    public static native int viewConfigId(long view, long[] idHigh, long[] idLow, long[] out1);
    public static native int viewRegisterJoiners(long view, byte[] hostBytes, int[] hostOff, int[] port);  // first id
    /** decideViewChange on the device (rapid_view_apply_cut): members in the cut leave, registered joiners in it are added;
     *  outOldToNew may be null.  RAPID_EUUID_SEEN (-4) = UUIDAlreadySeenException, nothing changed. */
    public static native int viewApplyCut(long view, int[] cutIds, int[] outOldToNew);
    /** identifiersSeen on the device: NodeIds of the members (index = node id) / of registered joiners */
    public static native int viewSetNodeIds(long view, long[] idHigh, long[] idLow);
    public static native int viewSetJoinerIds(long view, int firstJoinerId, long[] idHigh, long[] idLow);
    /** getCurrentConfigurationId from the device-resident identifiersSeen + ring 0 */
    public static native int viewCurrentConfigId(long view, long[] out1);

    // ---- MultiNodeCutDetector / alert-batch handler ----
    public static native long cdCreate(long view, int h, int l, long receivers, long receiverBegin, int modeFlags,
                                       long maxSubjects);
    public static native int cdDestroy(long cd);
    /** rapid_cd_apply_batch with direct buffers: dst int32[n], ring uint8[n], status uint8[n]; outputs may be null. */
    public static native int cdApplyBatch(long cd, long cfgId, long nCells, ByteBuffer dst, ByteBuffer ring,
                                          ByteBuffer status, ByteBuffer cellCfg, int deliveryFlags, ByteBuffer blocked,
                                          ByteBuffer bitmap, long permSeed, ByteBuffer outHash, ByteBuffer outHash2,
                                          ByteBuffer outLen, ByteBuffer outAnnounced);
    public static native int cdGetProposal(long cd, long receiver, int[] outIds);      // returns length or <0
    public static native int cdAggregate(long cd, int[] dst, byte[] ring, byte[] status, long receiver, int[] outIds);
    public static native int cdInvalidate(long cd, long receiver, int[] outIds);
    public static native int cdNumProposals(long cd, long receiver);
    public static native int cdClear(long cd);
    /** out4 = {sequences served in one pass, replayed batch by batch, receivers failing premise A1 / A2 in the last refusal} */
    public static native int cdSequenceStats(long cd, int[] out4);
    /** wait for asynchronous batches (wireApplyToDetector / fdetApplyToDetector with async = true); returns their latched status */
    public static native int cdSync(long cd);
    /** several BatchedAlertMessages in one call: batch b = cells [batchOff[b], batchOff[b+1]); announcedIn[r] = batch index or -1 */
    public static native int cdApplyBatches(long cd, long cfgId, int[] dst, byte[] ring, byte[] status, long[] cellCfg, long[] batchOff,
                                            ByteBuffer outHash, ByteBuffer outHash2, ByteBuffer outLen, ByteBuffer outAnnounced,
                                            ByteBuffer outAnnouncedIn);

    // ---- FastPaxos fast round ----
    public static native long fpCreate(long cfgId, long membershipSize, long senderCapacity, int device);
    public static native int fpDestroy(long fp);
    public static native int fpReset(long fp, long cfgId, long membershipSize);
    /** out6 = {decided, hashLo.. } see rapid_fp_tally; returns status */
    public static native int fpTally(long fp, int[] sender, long[] voteCfg, long[] hash, long[] hash2, int[] len,
                                     long[] out6);
    public static native int fpTallyCd(long fp, long cd, long comm, long[] out6);
    /** enqueue only (rapid_fp_tally_cd_async) / collect the last enqueued tally: out7 = out6 + {decidedInCall} */
    public static native int fpTallyCdAsync(long fp, long cd, long comm);
    public static native int fpResult(long fp, long[] out7);

    public static native long[] proposalFingerprint(int[] ids);

    // ---- classic Paxos fallback (Paxos.java) ----
    public static native long pxCreate(long cfgId, long membershipSize, long messageCapacity, int device);
    public static native int pxDestroy(long px);
    /** startPhase1a: returns 1 if crnd moved to (round, nodeIndex), 0 if ignored, <0 status */
    public static native int pxStartPhase1a(long px, int round, int nodeIndex);
    /** selectProposalUsingCoordinatorRule: index of the message whose vval is chosen, -1 = empty list, <-1 status-2 */
    public static native long pxCoordinatorRule(long px, int[] vrndRound, int[] vrndNode, long[] hash, long[] hash2, int[] len);
    /** handlePhase1bMessage over a batch; out6 = {proposed, triggerIndex, cvalHash, cvalHash2, cvalLen, nMessages} */
    public static native int pxPhase1b(long px, long[] msgCfg, int[] rndRound, int[] rndNode, int[] vrndRound, int[] vrndNode,
                                       long[] hash, long[] hash2, int[] len, long[] out6);
    /** handlePhase2bMessage over a batch; out5 = {decided, decidedIndex, hash, hash2, len} */
    public static native int pxPhase2b(long px, long[] msgCfg, int[] rndRound, int[] rndNode, int[] sender, long[] hash,
                                       long[] hash2, int[] len, long[] out5);
    public static native long pxaCreate(long cfgId, long nAcceptors, long acceptorBegin, int device);
    public static native int pxaDestroy(long pxa);
    public static native int pxaRegisterFastRoundVotesCd(long pxa, long cd);
    public static native long pxaPhase1a(long pxa, long msgCfg, int round, int nodeIndex);            // replies or <0
    public static native long pxaPhase2a(long pxa, long msgCfg, int round, int nodeIndex, long hash, long hash2, int len);
    public static native int pxPhase1bFromAcceptors(long px, long pxa, long permSeed, long[] out6);
    public static native int pxPhase2bFromAcceptors(long px, long pxa, long permSeed, long[] out5);

    // ---- wire-format ingest (rapid.proto bytes -> cells on the device) ----
    public static native long wireCreate(long view);
    /** only UP alerts of this configuration register joiners from now on */
    public static native int wireSetConfiguration(long wire, long cfgId);
    public static native int wireDestroy(long wire);
    /** bytes = BatchedAlertMessage.toByteArray() (or the RapidRequest, asRequest); out5 = {nMessages, nCells, nDropped, nNewJoiners, senderId} */
    public static native int wireDecodeAlerts(long wire, ByteBuffer bytes, int len, boolean asRequest, long[] out5);
    /** apply the cells of the last decode to a detector without leaving the device (rapid_wire_cells_dev + rapid_cd_apply_batch_dev) */
    public static native int wireApplyToDetector(long wire, long cd, long cfgId, long nCells);
    /** same, enqueue only (rapid_cd_apply_batch_dev_async): the status comes back from cdSync / fpTallyCd */
    public static native int wireApplyToDetectorAsync(long wire, long cd, long cfgId, long nCells);

    // ---- alert generation: the K PingPongFailureDetectors of every virtual node ----
    public static native long fdetCreate(long view, int failureThreshold, int bootstrapThreshold);
    public static native int fdetDestroy(long fdet);
    public static native int fdetReset(long fdet);
    /** one failure-detector interval; out2 = {nAlerts, nCells}; the cells stay on the device */
    public static native int fdetTick(long fdet, byte[] nodeFlags, byte[] edgeFail, long cfgId, long[] out2);
    public static native int fdetApplyToDetector(long fdet, long cd, long cfgId, long nCells);
}
