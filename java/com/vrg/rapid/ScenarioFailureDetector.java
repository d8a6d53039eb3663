/*
 * Seam 2 of INTEGRATION.md: the failure-injection half.  UNCOMPILED here (no JDK in the build image).
 *
 * implements IEdgeFailureDetectorFactory (monitoring/IEdgeFailureDetectorFactory.java:32-34) in the pattern of the
 * reference's test detector (src/test/java/com/vrg/rapid/StaticFailureDetector.java:26-62): the Runnable handed back for
 * a subject fires notifier.run() while the subject is on the scenario's failed list.  MembershipService schedules every
 * returned Runnable at failureDetectorIntervalInMs on its "msbg" thread (MembershipService.java:697-707) and re-creates
 * them after each view change (:433-434), so nothing here outlives a configuration.
 *
 * The same scenario also drives the DEVICE-side detectors of the virtual cluster (rapid_fdet_*: one PingPong-style
 * detector per (virtual node, ring) in HBM): tickVirtualCluster() advances all of them by one interval under the current
 * failed / blocked sets and feeds the alerts they raise straight into the cluster's cut detectors without leaving the GPU.
 * The real node (this JVM) and the N virtual ones therefore observe the same failures.
 */
package com.vrg.rapid;

import com.vrg.rapid.gpu.Native;
import com.vrg.rapid.monitoring.IEdgeFailureDetectorFactory;
import com.vrg.rapid.pb.Endpoint;

import java.util.Set;
import java.util.concurrent.ConcurrentHashMap;

final class ScenarioFailureDetector implements IEdgeFailureDetectorFactory {
    /** rapid_fdet_tick node flags (include/rapid_b200.h, RAPID_FD_*) */
    static final byte CRASHED = 1;
    static final byte INGRESS_BLOCKED = 2;
    static final byte EGRESS_BLOCKED = 4;
    static final byte BOOTSTRAPPING = 8;

    private final Set<Endpoint> failedNodes = ConcurrentHashMap.newKeySet();
    private final Object lock = new Object();
    private GpuMembershipView view;            // of the configuration the device detectors were created for
    private long fdet;                         // rapid_fdet handle, 0 = no virtual cluster attached
    private byte[] nodeFlags = new byte[0];

    /** The K detectors of THIS process: StaticFailureDetector.run(), one instance per subject. */
    @Override
    public Runnable createInstance(final Endpoint subject, final Runnable notifier) {
        return () -> {
            if (failedNodes.contains(subject)) {
                notifier.run();                // -> edgeFailureNotification, MembershipService.java:472-495
            }
        };
    }

    /** StaticFailureDetector.Factory.addFailedNodes: the subjects whose probes fail from now on. */
    void addFailedNodes(final Set<Endpoint> nodes) {
        failedNodes.addAll(nodes);
        synchronized (lock) {
            if (view != null) {
                for (final Endpoint e : nodes) {
                    final int id = view.tryIdOf(e);
                    if (id >= 0 && id < nodeFlags.length) {
                        nodeFlags[id] |= CRASHED;
                    }
                }
            }
        }
    }

    /** One-way partitions (the paper's Fig. 9 scenarios): the node's probes out, or the probes to it, are dropped. */
    void block(final Endpoint node, final boolean ingress, final boolean egress) {
        synchronized (lock) {
            if (view == null) {
                return;
            }
            final int id = view.tryIdOf(node);
            if (id >= 0 && id < nodeFlags.length) {
                nodeFlags[id] |= (byte) ((ingress ? INGRESS_BLOCKED : 0) | (egress ? EGRESS_BLOCKED : 0));
            }
        }
    }

    /**
     * Attach the virtual cluster of a configuration: every virtual node gets its K device-resident detectors
     * (PingPongFailureDetector.java:38-121 semantics: threshold 10, bootstrap tolerance 30).  Called again after a view
     * change, like MembershipService.createFailureDetectorsForCurrentConfiguration (:697-707).
     */
    void attach(final GpuMembershipView newView, final int members) {
        synchronized (lock) {
            if (fdet != 0) {
                Native.fdetDestroy(fdet);
            }
            view = newView;
            nodeFlags = new byte[members];
            for (final Endpoint e : failedNodes) {
                final int id = newView.tryIdOf(e);
                if (id >= 0 && id < members) {
                    nodeFlags[id] |= CRASHED;
                }
            }
            fdet = Native.fdetCreate(newView.handle(), 10, 30);
            if (fdet == 0) {
                throw new IllegalStateException(Native.lastError());
            }
        }
    }

    /**
     * One failure-detector interval of all N x K virtual detectors; the alerts they raise are applied to the cluster's cut
     * detectors on the device (no host copy of the cells).  Returns the number of alert cells of the interval.
     */
    long tickVirtualCluster(final long cutDetector, final long configurationId) {
        synchronized (lock) {
            if (fdet == 0) {
                return 0;
            }
            final long[] out = new long[2];                   // {alerts, cells}
            int rc = Native.fdetTick(fdet, nodeFlags, null, configurationId, out);
            if (rc == 0 && out[1] > 0) {
                rc = Native.fdetApplyToDetector(fdet, cutDetector, configurationId, out[1]);
            }
            if (rc != 0) {
                throw new IllegalStateException(Native.lastError());
            }
            return out[1];
        }
    }

    void shutdown() {
        synchronized (lock) {
            if (fdet != 0) {
                Native.fdetDestroy(fdet);
                fdet = 0;
            }
        }
    }
}
