/*
 * Seam 2 of INTEGRATION.md: a virtual cluster behind Rapid's messaging SPI.  UNCOMPILED here (no JDK).
 *
 * implements IMessagingClient (messaging/IMessagingClient.java:25-49) and IMessagingServer (:24-41): instead of a socket,
 * sendMessageBestEffort(remote, BATCHEDALERTMESSAGE) applies the batch to the HBM-resident detectors of every virtual
 * node (UnicastToAllBroadcaster.java:46-52 sends the same request to all members: the first unicast of a broadcast
 * triggers the device call, the rest are no-ops), and FASTROUNDPHASE2BMESSAGE votes are tallied on the device.  One mutex
 * serialises callers (protocol thread / "msbg" batcher thread, MembershipService.java:630).
 */
package com.vrg.rapid;

import com.google.common.util.concurrent.Futures;
import com.google.common.util.concurrent.ListenableFuture;
import com.vrg.rapid.gpu.Native;
import com.vrg.rapid.messaging.IMessagingClient;
import com.vrg.rapid.messaging.IMessagingServer;
import com.vrg.rapid.pb.AlertMessage;
import com.vrg.rapid.pb.BatchedAlertMessage;
import com.vrg.rapid.pb.Endpoint;
import com.vrg.rapid.pb.RapidRequest;
import com.vrg.rapid.pb.RapidResponse;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

final class GpuSimMessaging implements IMessagingClient, IMessagingServer {
    private final Object lock = new Object();
    private final GpuMembershipView view;
    private final long cd;
    private final long fp;
    private final long configurationId;
    private BatchedAlertMessage lastApplied;      // identity of the broadcast already applied

    GpuSimMessaging(final GpuMembershipView view, final long configurationId, final int H, final int L,
                    final int members) {
        this.view = view;
        this.configurationId = configurationId;
        this.cd = Native.cdCreate(view.handle(), H, L, members, 0, 0 /* SERVICE, bucketed */, 0);
        this.fp = Native.fpCreate(configurationId, members, members, 0);
    }

    @Override
    public ListenableFuture<RapidResponse> sendMessageBestEffort(final Endpoint remote, final RapidRequest msg) {
        synchronized (lock) {
            switch (msg.getContentCase()) {
                case BATCHEDALERTMESSAGE:
                    if (msg.getBatchedAlertMessage() != lastApplied) {   // one device call per broadcast
                        lastApplied = msg.getBatchedAlertMessage();
                        applyBatch(lastApplied);
                        final long[] out = new long[6];
                        Native.fpTallyCd(fp, cd, 0, out);                // every virtual node that announced votes
                    }
                    break;
                default:
                    break;                                               // probes, joins: not simulated on the device
            }
        }
        return Futures.immediateFuture(RapidResponse.getDefaultInstance());
    }

    private void applyBatch(final BatchedAlertMessage batch) {
        int cells = 0;
        for (final AlertMessage m : batch.getMessagesList()) {
            cells += m.getRingNumberCount();
        }
        final ByteBuffer dst = ByteBuffer.allocateDirect(4 * cells).order(ByteOrder.nativeOrder());
        final ByteBuffer ring = ByteBuffer.allocateDirect(cells);
        final ByteBuffer status = ByteBuffer.allocateDirect(cells);
        final ByteBuffer cfg = ByteBuffer.allocateDirect(8 * cells).order(ByteOrder.nativeOrder());
        for (final AlertMessage m : batch.getMessagesList()) {
            final int id = view.idOf(m.getEdgeDst(), true);
            for (int i = 0; i < m.getRingNumberCount(); i++) {       // one cell per ring number (MultiNodeCutDetector.java:79-80)
                dst.putInt(id);
                ring.put((byte) m.getRingNumber(i));
                status.put((byte) m.getEdgeStatusValue());
                cfg.putLong(m.getConfigurationId());
            }
        }
        Native.cdApplyBatch(cd, configurationId, cells, dst, ring, status, cfg, 0, null, null, 0L, null, null, null, null);
    }

    @Override
    public ListenableFuture<RapidResponse> sendMessage(final Endpoint remote, final RapidRequest msg) {
        return sendMessageBestEffort(remote, msg);
    }

    @Override
    public void start() {
    }

    @Override
    public void shutdown() {
        Native.cdDestroy(cd);
        Native.fpDestroy(fp);
    }

    @Override
    public void setMembershipService(final MembershipService service) {
    }
}
