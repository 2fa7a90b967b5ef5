/*
 * Serialises the matched plan fragment into the off-heap sd_plan_desc block the C ABI takes, and moves
 * buffers/rows across the JNI boundary.  NOT COMPILED in this repository's container (no JDK/Scala); a sketch
 * against the reference's class names for a maintainer to adapt (see INTEGRATION.md).
 */
package io.snappydata.gpu

import java.nio.ByteBuffer

import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.catalyst.expressions._
import org.apache.spark.sql.catalyst.expressions.aggregate._
import org.apache.spark.sql.execution.{FilterExec, ProjectExec, SparkPlan}
import org.apache.spark.sql.execution.aggregate.SnappyHashAggregateExec
import org.apache.spark.sql.execution.columnar.{ColumnBatchIterator, ColumnTableScan}
import org.apache.spark.sql.execution.metric.SQLMetric
import org.apache.spark.sql.types._
import org.apache.spark.unsafe.Platform

/** Off-heap image of sd_plan_desc (include/snappy_gpu.h) plus what the operator needs at run time. */
final class GpuPlanDesc(val address: Long, val numLiterals: Int, val scan: ColumnTableScan,
    val literals: Seq[Expression /* ParamLiteral | TokenLiteral | Literal */], val outputSchema: StructType) {
  /** evaluates the tokenised constants of THIS execution (ParamLiteral.value) into sd_literal[] */
  def literalsAddress(): Long = GpuPlanSerializer.writeLiterals(literals)
}

object GpuPlanSerializer {
  def numDevices: Int = java.lang.Integer.getInteger("snappydata.gpu.devices", 1)

  private def sdType(dt: DataType): Option[Int] = dt match {
    case BooleanType => Some(1); case ByteType => Some(2); case ShortType => Some(3); case IntegerType => Some(4)
    case LongType => Some(5); case FloatType => Some(6); case DoubleType => Some(7); case DateType => Some(8)
    case TimestampType => Some(9); case StringType => Some(10)
    case d: DecimalType if d.precision <= 18 => Some(11)
    case _ => None
  }

  /** Some(desc) iff the fragment is SnappyHashAggregateExec(Partial) over [Project] [Filter] ColumnTableScan with
    * expressions the C ABI can express (sd_op in include/snappy_gpu.h); anything else stays on the stock operators. */
  def tryBuild(agg: SnappyHashAggregateExec): Option[GpuPlanDesc] = {
    def unwrap(p: SparkPlan, filters: Seq[Expression]): Option[(ColumnTableScan, Seq[Expression])] = p match {
      case FilterExec(cond, child) => unwrap(child, filters :+ cond)
      case ProjectExec(_, child) => unwrap(child, filters) // projections are inlined into the expression trees
      case scan: ColumnTableScan => Some((scan, filters))
      case _ => None
    }
    unwrap(agg.child, Nil).flatMap { case (scan, filters) =>
      // flatten: scan.output -> sd_column[]; filters.reduce(And) / grouping / aggregate children -> sd_expr[]
      // (Cast(l_shipdate, StringType) >= '1994-01-01' produced by PromoteStrings is normalised to an int-day compare,
      //  SURVEY.md Appendix B.8); literals become slots in evaluation order
      ??? // elided: mechanical tree walk writing the structs of include/snappy_gpu.h with Platform.putInt/putLong
    }
  }

  def writeLiterals(literals: Seq[Expression]): Long = ???

  /** per batch: value buffers via getColumnLob, deltas via the iterator's delta lookups (ColumnBatchIterator.scala:122-163) */
  def submitBatch(plan: Long, it: ColumnBatchIterator, stats: ByteBuffer, desc: GpuPlanDesc): Unit = {
    val cols = desc.scan.output                              // scan columns in plan order
    val nCols = cols.length
    val nTableCols = desc.scan.relationSchema.length
    val statsRow = org.apache.spark.sql.collection.SharedUtils.toUnsafeRow(stats, 1 + 3 * nTableCols)
    // count < 0 marks a batch that carries delta updates (ColumnTableScan.scala:524-528); the engine takes the
    // row count from here and never reads the count field of the stats row
    val numRows = math.abs(statsRow.getInt(0))
    // an old-format delta stats row means the full stats may be stale: no skipping for this batch (:536-539)
    val passStats = it.getCurrentDeltaStats == null
    val addrs = new Array[Long](nCols); val lens = new Array[Long](nCols); val heap = new Array[Array[Byte]](nCols)
    val d0a = new Array[Long](nCols); val d0l = new Array[Long](nCols)
    val d1a = new Array[Long](nCols); val d1l = new Array[Long](nCols)
    def place(buf: ByteBuffer, i: Int, a: Array[Long], l: Array[Long], h: Array[Array[Byte]]): Unit = {
      l(i) = buf.remaining()
      if (buf.isDirect) a(i) = io.snappydata.gpu.DirectBuffers.address(buf) + buf.position() // GetDirectBufferAddress
      else { a(i) = 0L; h(i) = buf.array() /* arrayOffset + position == 0 for store buffers (ColumnTableScan.scala:430-437) */ }
    }
    var i = 0
    while (i < nCols) {
      val tableCol = desc.scan.baseRelation.schema.fieldIndex(cols(i).name) + 1 // ColumnFormatKey.columnIndex is 1-based
      place(it.getColumnLob(tableCol - 1), i, addrs, lens, heap)
      // deltas are always direct/off-heap copies owned by the iterator (ColumnBatchIterator.scala:122-150)
      val u0 = it.getUpdatedColumnBuffer(tableCol, 0); if (u0 ne null) place(u0, i, d0a, d0l, null)
      val u1 = it.getUpdatedColumnBuffer(tableCol, 1); if (u1 ne null) place(u1, i, d1a, d1l, null)
      i += 1
    }
    val del = it.getDeletedColumnBuffer
    val (delAddr, delLen) = if (del eq null) (0L, 0L) else (DirectBuffers.address(del) + del.position(), del.remaining().toLong)
    val (stAddr, stLen) = if (passStats) (DirectBuffers.addressOrCopy(stats), stats.remaining().toLong) else (0L, 0L)
    // throws RuntimeException(sd_last_error()) on a non-zero status; the buffers may be released when it returns
    SnappyGpuNative.batchSubmit(plan, numRows, nCols, addrs, lens, heap, d0a, d0l, d1a, d1l, delAddr, delLen,
      stAddr, stLen, nTableCols, it.getCurrentBucketId, it.getCurrentBatchId)
  }

  /** un-rolled-over rows as UnsafeRows (ColumnTableScan.scala:572-588) */
  def submitRowBuffer(plan: Long, rows: Iterator[InternalRow], desc: GpuPlanDesc): Unit = ???

  /** sd_plan_finish -> Iterator[UnsafeRow] over the returned [int64 size][row] stream (UnsafeRow.pointTo) */
  def finishToUnsafeRows(plan: Long, desc: GpuPlanDesc, numOutputRows: SQLMetric): Iterator[InternalRow] = ???
}
